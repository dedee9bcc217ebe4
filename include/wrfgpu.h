/* wrfgpu.h -- C-ABI of the B200-native WRF-ARW dry dynamical-core step.
 *
 * This is the drop-in boundary for the reference's
 *     SUBROUTINE solve_em ( grid , config_flags , <dummy_new_args.inc> )
 * (dyn_em/solve_em.F:3-7, called from share/solve_interface.F:50-55).  The
 * Fortran host binds these symbols with ISO_C_BINDING (see
 * wrf_b200/fortran/module_wrfgpu.F90 and INTEGRATION.md); every signature
 * uses plain pointers, ints and floats only.
 *
 * Conventions
 *   - all arrays are real*4, WRF "XZY" memory order: 3-D (ims:ime,kms:kme,jms:jme)
 *     i fastest, 2-D (ims:ime,jms:jme), 1-D (kms:kme)
 *     (frame/module_driver_constants.F:15, dyn_em/module_em.F:68);
 *   - host arrays stay owned by the caller (Fortran `grid%...`); the library
 *     owns the device mirrors and never frees or reallocates host memory;
 *   - every entry point returns 0 on success, non-zero on error; the message
 *     is available from wrfgpu_last_error() and the Fortran shim turns it into
 *     CALL wrf_error_fatal(msg) (frame/module_wrf_error.F);
 *   - unsupported namelist values fail loudly at wrfgpu_create -- there is no
 *     CPU fallback.
 */
#ifndef WRFGPU_H
#define WRFGPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Index triplets of one domain on one rank: domain (ids..kde), memory
 * (ims..kme) and patch (ips..kpe) dims, as returned by get_ijk_from_grid
 * (dyn_em/solve_em.F:288-296).  The GPU build runs numtiles=1, so the tile
 * (its..jte) equals the patch (SURVEY 8b "Threading"). */
typedef struct wrfgpu_dims {
  int ids, ide, jds, jde, kds, kde;
  int ims, ime, jms, jme, kms, kme;
  int ips, ipe, jps, jpe, kps, kpe;
} wrfgpu_dims;

/* The column of namelist.input / grid scalars the path reads
 * (grid_config_rec_type, generated from Registry/Registry.EM_COMMON rconfig
 * rows; flattened to PODs because config_flags cannot cross a C ABI). */
typedef struct wrfgpu_config {
  /* &bdy_control */
  int periodic_x, periodic_y;
  int symmetric_xs, symmetric_xe, symmetric_ys, symmetric_ye;
  int open_xs, open_xe, open_ys, open_ye;
  int specified, nested, polar;          /* specified / nested: the lateral boundary zone is driven by the boundary arrays
                                            (the zone widths are at the end of this struct); polar must be 0
                                            (polar filters are out of scope; rejected at create) */
  /* &dynamics */
  int rk_ord;                            /* 3 (2 also accepted) */
  int time_step_sound;                   /* 0 => derived, solve_em.F:446-462 */
  int h_mom_adv_order, v_mom_adv_order;  /* 5 / 3 */
  int h_sca_adv_order, v_sca_adv_order;  /* 5 / 3 */
  int moist_adv_opt, scalar_adv_opt;     /* 0 = original, 1 = positive-definite, 2 = monotonic,
                                            3 = WENO, 4 = WENO + positive-definite
                                            (Registry.EM_COMMON:3275-3279) */
  int momentum_adv_opt;                  /* 1 = original, 3 = WENO (Registry.EM_COMMON:3272-3273) */
  int non_hydrostatic;                   /* 1 */
  int top_lid;                           /* 0/1 */
  int damp_opt;                          /* 0, 2 (rk_rayleigh_damp), 3 (implicit w) */
  int w_damping;                         /* 0/1 */
  int pert_coriolis;                     /* 0/1 */
  int use_theta_m;                       /* 0/1 */
  int hypsometric_opt;                   /* 1/2 */
  int phi_adv_z;                         /* 1/2 */
  int zadvect_implicit;                  /* > 0: IEVA on the last RK substep (module_ieva_em.F) */
  int diff_opt, km_opt;                  /* diff_opt 0, 1 (coordinate surfaces), 2 (physical space);
                                            km_opt 1 (constant K), 2 (1.5-order TKE; needs diff_opt 2), 3 (3-D
                                            Smagorinsky) or 4 (2-D Smagorinsky) */
  int diff_6th_opt;                      /* 0, 1 or 2 (sixth_order_diffusion; 2 = no up-gradient fluxes) */
  int mp_physics;                        /* 0 = none(passive qv), 1 = Kessler */
  int time_step;                         /* integer namelist time_step; only its
                                            sign is used (module_advect_em.F:204) */
  int map_proj;                          /* 0 for ideal cases; 6 rejected */
  float dt;                              /* grid%dt */
  float dx, dy;                          /* rdx = 1/dx, rdy = 1/dy */
  float epssm, smdiv, emdiv;
  float dampcoef, zdamp;
  float w_crit_cfl;
  float p_top;
  float cf1, cf2, cf3, cfn, cfn1;        /* Registry.EM_COMMON:405-407, 358-359 */
  float khdif, kvdif;
  /* species counts: moist = qv[,qc,qr]; scalar = generic passive species
   * advected like scalar/tracer/chem (dyn_em/solve_em.F:2200-2955) */
  int num_moist;                         /* 0..3 (without the WRF dummy slot 1) */
  int num_scalar;                        /* 0..8 */
  /* RSL_LITE process mesh (external/RSL_LITE/module_dm.F:98-175) */
  int ntasks_x, ntasks_y, mytask;        /* mytask = mytask_y*ntasks_x + mytask_x */
  /* 6th-order horizontal filter (Registry.EM_COMMON:2860-2863, 2893-2896) */
  float diff_6th_factor;                 /* 0.12 */
  float diff_6th_thresh;                 /* 0.10 */
  int diff_6th_slopeopt;                 /* 0/1: taper the filter over steep terrain */
  int moist_mix6_off, scalar_mix6_off;   /* 1 = no 6th-order filter on that family */
  /* eddy-viscosity closures (Registry.EM_COMMON rconfig c_s, c_k, mix_full_fields) */
  float c_s;                             /* Smagorinsky coefficient, 0.25 (km_opt = 3, 4) */
  float c_k;                             /* TKE coefficient, 0.15 (km_opt = 2) */
  int mix_full_fields;                   /* 0/1: vertical shear of the full u, v in defor13/23 instead of u - u_base */
  /* surface fluxes of the physical-space vertical diffusion (diff_opt = 2; module_diffusion_em.F:4160-4353) */
  int isfflx;                            /* 0: drag and heat flux from the two constants below; 1, 2: from ustm / hfx / qfx */
  float tke_drag_coefficient;            /* 0. */
  float tke_heat_flux;                   /* 0. (K m/s) */
  /* 3-D Smagorinsky closure (km_opt = 3; smag_km, module_diffusion_em.F:1777) */
  int mix_isotropic;                     /* 0 = separate horizontal / vertical length scales, 1 = (dx dy dz)^(1/3) */
  float mix_upper_bound;                 /* 0.1: K <= mix_upper_bound * l^2 / dt */
  /* 1.5-order TKE closure (km_opt = 2 with diff_opt = 2): the prognostic tke_1 / tke_2 (Registry.EM_COMMON:312) */
  int tke_adv_opt;                       /* as scalar_adv_opt (Registry.EM_COMMON:2884) */
  float tke_upper_bound;                 /* 1000. (bound_tke, module_em.F:2616) */
  /* w_damp's CFL maxima (module_big_step_utilities_em.F:2503-2711) are folded into grid%max_vert_cfl /
   * max_horiz_cfl only with the adaptive time step (solve_em.F:914-930); 1 = compute them on the device */
  int use_adaptive_time_step;
  /* specified / nested lateral boundaries (&bdy_control; share/module_bc.F:1122-2693, dyn_em/module_bc_em.F:17-700) */
  int spec_bdy_width;                    /* 5: rows held by the boundary arrays */
  int spec_zone;                         /* 1: outer rows set from the boundary file */
  int relax_zone;                        /* 4: rows relaxed towards it (spec_zone + relax_zone = spec_bdy_width) */
  int have_bcs_moist, have_bcs_scalar;   /* 0/1: every moist species / scalar has boundary arrays (always so when nested) */
} wrfgpu_config;

typedef struct wrfgpu_solver wrfgpu_solver;   /* opaque */

/* -- life cycle ---------------------------------------------------------- */
/* replaces alloc_and_configure_domain's role for the device mirrors
 * (main/module_wrf_top.F:361) */
int  wrfgpu_create (const wrfgpu_dims *d, const wrfgpu_config *c, int device,
                    wrfgpu_solver **out);
int  wrfgpu_destroy(wrfgpu_solver *s);
const char *wrfgpu_last_error(void);
int  wrfgpu_version(void);
/* CUDA devices visible to this process (the Fortran shim maps rank -> device with it); 0 when there is none */
int  wrfgpu_device_count(void);

/* -- state transfer ------------------------------------------------------ */
/* Registry names without the grid% prefix: "u_2","t_1","mu_2","phb","fnm",
 * "moist_qv","moist_old_qv","scalar_1",...  (Registry/Registry.EM_COMMON:158-457).
 * n = number of floats in the host buffer; must match the field's full memory
 * extent.  wrfgpu_field_size returns that extent (0 if the name is unknown). */
size_t wrfgpu_field_size(const wrfgpu_solver *s, const char *name);
int  wrfgpu_set_field(wrfgpu_solver *s, const char *name, const float *host, size_t n);
int  wrfgpu_get_field(wrfgpu_solver *s, const char *name, float *host, size_t n);
/* pointer-table forms used by the Fortran shim (one call per solve_em) */
int  wrfgpu_upload  (wrfgpu_solver *s, int nfields, const char *const *names,
                     const float *const *host);
int  wrfgpu_download(wrfgpu_solver *s, int nfields, const char *const *names,
                     float *const *host);
int  wrfgpu_num_fields(const wrfgpu_solver *s);
const char *wrfgpu_field_name(const wrfgpu_solver *s, int idx);
void *wrfgpu_field_devptr(wrfgpu_solver *s, const char *name);

/* -- history / restart snapshots --------------------------- */
/* replaces the contract "the host grid% arrays are current when
 * med_before_solve_io runs" (frame/module_integrate.F:406;
 * share/mediation_integrate.F: med_before_solve_io -> med_hist_out,
 * med_restart_out at the history / restart alarms).  _begin copies the listed
 * fields device-to-device into staging buffers on the compute stream and
 * returns; the device-to-host copies into pinned memory overlap the following
 * time steps on a second stream.  _wait blocks until they have landed; _host
 * then returns the pinned buffer of one field (valid until the next _begin;
 * NULL if the field is not part of the current snapshot).  _busy: 1 while the
 * copies are in flight, 0 when done, -1 on error. */
int  wrfgpu_snapshot_begin(wrfgpu_solver *s, int nfields, const char *const *names);
int  wrfgpu_snapshot_wait (wrfgpu_solver *s);
int  wrfgpu_snapshot_busy (wrfgpu_solver *s);
const float *wrfgpu_snapshot_host(wrfgpu_solver *s, const char *name, size_t *n);

/* -- the hot path -------------------------------------------------------- */
/* one model time step of one domain = dyn_em/solve_em.F:578-4848 with the
 * state resident in HBM */
int  wrfgpu_step(wrfgpu_solver *s, int nsteps);
/* strict drop-in form: upload the prognostic set from host grid% arrays, run
 * one solve_em, download it again (what wrf.exe sees when I/O needs host data
 * every step).  `names/host` list the fields to move in each direction. */
int  wrfgpu_solve_em(wrfgpu_solver *s,
                     int n_in,  const char *const *in_names,  const float *const *in_host,
                     int n_out, const char *const *out_names, float *const *out_host);
int  wrfgpu_sync(wrfgpu_solver *s);
/* specified / nested lateral boundaries.  The boundary arrays are fields like any other: "<f>_bxs", "<f>_bxe",
 * "<f>_bys", "<f>_bye" and the tendencies "<f>_btxs" ... "<f>_btye" for f = u, v, w, ph, t, mu, moist_qv (and every moist
 * species / scalar_N that has them), in the reference's layout (jms:jme | ims:ime, kds:kde, spec_bdy_width), mu with one
 * level (Registry.EM_COMMON 'b' entries); "fcx", "gcx" are the relaxation weights start_em computes (lbc_fcx_gcx,
 * dyn_em/module_bc_em.F:1297).  grid%dtbc, the time since the arrays were valid, is advanced by every step
 * (solve_em.F:377); the host sets it back to 0 when it uploads the next boundary time (share/input_wrf.F:1834,
 * mediation_integrate.F:1529). */
int  wrfgpu_set_dtbc(wrfgpu_solver *s, float dtbc);
int  wrfgpu_get_dtbc(wrfgpu_solver *s, float *dtbc);
/* run on the caller's CUDA stream (cudaStream_t) instead of the solver's own */
int  wrfgpu_set_stream(wrfgpu_solver *s, void *cuda_stream);
/* per-kernel cudaEvent statistics collected while phase timing is on, as text
 * lines "kernel count total_ms"; returns the buffer size needed */
size_t wrfgpu_kernel_report(wrfgpu_solver *s, char *buf, size_t cap);
/* grid%max_vert_cfl / grid%max_horiz_cfl (solve_em.F:914-930): the running maxima of w_damp's Courant numbers over
 * every RK substep since the last reset, over all ranks of the mesh (the wrf_dm_maxval of adapt_timestep,
 * dyn_em/adapt_timestep_em.F).  Needs use_adaptive_time_step = 1.  reset != 0 zeroes them after the read. */
int  wrfgpu_max_cfl(wrfgpu_solver *s, float *max_vert_cfl, float *max_horiz_cfl, int reset);
/* number of kernels this library launched since create (bench "gpu_launches") */
long long wrfgpu_launch_count(const wrfgpu_solver *s);
/* cudaEvent timing of the kernels launched by phase `name` since the last
 * reset, in ms (0 if unknown); reset with name == NULL */
double wrfgpu_phase_ms(wrfgpu_solver *s, const char *name);

/* -- per-routine entry points (kernel-level parity tests) --
 * Each runs on the resident state, same semantics as the Fortran routine named
 * in the comment with its=ips.. tile bounds. */
int wrfgpu_start_domain     (wrfgpu_solver *s);               /* start_em.F: BC fills before step 1 */
int wrfgpu_rk_step_prep     (wrfgpu_solver *s, int rk_step);  /* module_em.F:37            */
int wrfgpu_rk_phys_bc_dry_1 (wrfgpu_solver *s);               /* solve_em.F:726-754        */
int wrfgpu_init_zero_tendency(wrfgpu_solver *s);              /* module_em.F:1920          */
int wrfgpu_diffusion_prep   (wrfgpu_solver *s);               /* module_first_rk_step_part2.F:427-680:
                                                                 compute_diff_metrics, cal_deform_and_div,
                                                                 calculate_km_kh (km_opt 4), phy_bc */
int wrfgpu_rk_tendency      (wrfgpu_solver *s, int rk_step);  /* module_em.F:190 + :959    */
int wrfgpu_small_step_prep  (wrfgpu_solver *s, int rk_step);  /* module_small_step_em.F:16,
                                                                 calc_p_rho(0) :438, calc_coef_w :570,
                                                                 bc fills solve_em.F:1186-1263 */
int wrfgpu_advance_uv       (wrfgpu_solver *s, int rk_step);  /* module_small_step_em.F:654 */
int wrfgpu_advance_mu_t     (wrfgpu_solver *s, int rk_step);  /* :969  */
int wrfgpu_advance_w        (wrfgpu_solver *s, int rk_step);  /* :1178 */
int wrfgpu_sumflux          (wrfgpu_solver *s, int rk_step, int iteration); /* :1473 */
int wrfgpu_calc_p_rho       (wrfgpu_solver *s, int iteration);/* :438 + bc fills solve_em.F:1670-1713 */
int wrfgpu_small_step_finish(wrfgpu_solver *s, int rk_step);  /* calc_mu_uv_1 + :295 */
int wrfgpu_rk_scalar_step   (wrfgpu_solver *s, int rk_step);  /* solve_em.F:1850-2955: update_scalar_pd,
                                                                 rk_scalar_tend, rk_update_scalar */
int wrfgpu_calc_p_rho_phi   (wrfgpu_solver *s);               /* module_big_step_utilities_em.F:953 */
int wrfgpu_rk_phys_bc_dry_2 (wrfgpu_solver *s);               /* solve_em.F:3245-3343 */
int wrfgpu_microphysics     (wrfgpu_solver *s);               /* solve_em.F:3629-4206 (Kessler) */
int wrfgpu_end_of_step      (wrfgpu_solver *s);               /* solve_em.F:4437-4759 */
/* one physical-BC fill, share/module_bc.F:202 (2-D) / :651 (3-D); `var` is the
 * reference's one-letter stagger code */
int wrfgpu_set_physical_bc  (wrfgpu_solver *s, const char *name, char var);

/* -- decomposition + halo exchange (external/RSL_LITE) -------------------- */
/* module_dm.F:98-175 mpaspect/compute_mesh */
int wrfgpu_compute_mesh(int ntasks, int nproc_x, int nproc_y, int *ntasks_x, int *ntasks_y);
/* task_for_point.c:6-14 + module_dm.F:924-1003: patch + memory dims of `mytask` */
int wrfgpu_patch_dims(int ids, int ide, int jds, int jde, int kds, int kde,
                      int ntasks_x, int ntasks_y, int mytask, wrfgpu_dims *out);
/* Ranks as processes (one per GPU: wrf.exe under mpirun, bench.py under torchrun): NCCL over
 * NVLink.  The 128-byte unique id is created on rank 0 and broadcast by the host (MPI_Bcast in
 * wrf.exe where wrf_dm_initialize builds its communicator, external/RSL_LITE/module_dm.F:177-232;
 * torch.distributed in bench.py). */
int wrfgpu_comm_unique_id(void *id128);
int wrfgpu_comm_init(wrfgpu_solver *s, const void *id128, int rank, int nranks);
/* Ranks as solvers of ONE process (one host thread each, on one or several GPUs with peer
 * access): neighbours read each other's send buffers directly; `all` is in rank order. */
int wrfgpu_comm_init_local(wrfgpu_solver *const *all, int n);
/* run one named Registry halo ("HALO_EM_A", "HALO_EM_D2_5", ...) on the resident state; a no-op
 * on a 1x1 mesh exactly as the generated HALO_*.inc are (Registry.EM_COMMON:3418-3551).  A
 * HALO_EM_x and the PERIOD_* include that follows it in solve_em.F are one exchange. */
int wrfgpu_halo(wrfgpu_solver *s, const char *halo_name);
/* host logic behind it, callable without a GPU:
 * the width and the comma-separated Registry field names of an exchange for this species set */
int wrfgpu_halo_spec(const wrfgpu_config *c, const char *halo_name, int *width, char *fields, size_t cap);
/* the index box this rank packs (send=1) or fills (send=0) in direction dir (0=-y 1=+y: Y phase,
 * 2=-x 3=+x: X phase, run after Y so corners propagate) and the peer rank; returns 0 when the rank
 * has no neighbour there (external/RSL_LITE/c_code.c:405-700, period.c, module_dm.F:2761) */
int wrfgpu_halo_box(const wrfgpu_dims *d, const wrfgpu_config *c, int width, int dir, int send,
                    int *peer, int *is, int *ie, int *js, int *je);
/* A HALO_EM_x merged with its PERIOD_* twin carries the PERIOD width (e.g. HALO_EM_A 1, PERIOD_BDY_EM_A 2): a neighbour
 * inside the domain only gets the Registry's HALO width, which this returns (Registry.EM_COMMON:3425, :3447-3450);
 * _box2 is wrfgpu_halo_box with that width for links that do not cross a periodic edge (0 = the same as `width`). */
int wrfgpu_halo_regular_width(const char *halo_name, int width);
int wrfgpu_halo_box2(const wrfgpu_dims *d, const wrfgpu_config *c, int width, int regular_width, int dir, int send,
                     int *peer, int *is, int *ie, int *js, int *je);
/* bytes this rank has sent in halo exchanges since create (bench: NVLink traffic) */
long long wrfgpu_halo_bytes(const wrfgpu_solver *s);

#ifdef __cplusplus
}
#endif
#endif /* WRFGPU_H */
