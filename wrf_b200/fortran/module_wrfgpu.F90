!WRF:MODEL_LAYER:DYNAMICS
!
! module_wrfgpu -- ISO_C_BINDING interface to libwrfgpu.so (include/wrfgpu.h), the B200-native
! implementation of the dyn_em solve_em time step.  This is the thin host layer the north star
! asks for: wrf.exe stays Fortran, the Registry-defined state stays in TYPE(domain), and this
! module hands the addresses of grid%u_2, grid%t_2, ... to the C-ABI.  No Fortran compiler exists
! in the image this was developed in, so the module has been desk-checked against
! include/wrfgpu.h (struct member order, by-value vs by-reference; tests/test_abi.py also compares the
! member lists of the two BIND(C) types below with the header) but not compiled here.
!
MODULE module_wrfgpu
   USE, INTRINSIC :: iso_c_binding
   IMPLICIT NONE
   PRIVATE
   PUBLIC :: wrfgpu_dims, wrfgpu_config, wrfgpu_check
   PUBLIC :: wrfgpu_create, wrfgpu_destroy, wrfgpu_last_error, wrfgpu_upload, wrfgpu_download, wrfgpu_step, &
             wrfgpu_solve_em, wrfgpu_start_domain, wrfgpu_sync, wrfgpu_comm_unique_id, wrfgpu_comm_init, &
             wrfgpu_compute_mesh, wrfgpu_patch_dims, wrfgpu_field_size, wrfgpu_device_count, wrfgpu_max_cfl, wrfgpu_set_dtbc, &
             wrfgpu_snapshot_begin, wrfgpu_snapshot_wait, wrfgpu_snapshot_busy, wrfgpu_snapshot_host

   ! wrfgpu_dims: the three index triplets of get_ijk_from_grid (dyn_em/solve_em.F:288-296)
   TYPE, BIND(C) :: wrfgpu_dims
      INTEGER(c_int) :: ids, ide, jds, jde, kds, kde
      INTEGER(c_int) :: ims, ime, jms, jme, kms, kme
      INTEGER(c_int) :: ips, ipe, jps, jpe, kps, kpe
   END TYPE wrfgpu_dims

   ! wrfgpu_config: the namelist column of this domain, flattened (grid_config_rec_type cannot cross a C ABI)
   TYPE, BIND(C) :: wrfgpu_config
      INTEGER(c_int) :: periodic_x, periodic_y
      INTEGER(c_int) :: symmetric_xs, symmetric_xe, symmetric_ys, symmetric_ye
      INTEGER(c_int) :: open_xs, open_xe, open_ys, open_ye
      INTEGER(c_int) :: specified, nested, polar
      INTEGER(c_int) :: rk_ord, time_step_sound
      INTEGER(c_int) :: h_mom_adv_order, v_mom_adv_order, h_sca_adv_order, v_sca_adv_order
      INTEGER(c_int) :: moist_adv_opt, scalar_adv_opt, momentum_adv_opt
      INTEGER(c_int) :: non_hydrostatic, top_lid, damp_opt, w_damping, pert_coriolis
      INTEGER(c_int) :: use_theta_m, hypsometric_opt, phi_adv_z, zadvect_implicit
      INTEGER(c_int) :: diff_opt, km_opt, diff_6th_opt, mp_physics, time_step, map_proj
      REAL(c_float)  :: dt, dx, dy, epssm, smdiv, emdiv, dampcoef, zdamp, w_crit_cfl
      REAL(c_float)  :: p_top, cf1, cf2, cf3, cfn, cfn1, khdif, kvdif
      INTEGER(c_int) :: num_moist, num_scalar, ntasks_x, ntasks_y, mytask
      REAL(c_float)  :: diff_6th_factor, diff_6th_thresh
      INTEGER(c_int) :: diff_6th_slopeopt, moist_mix6_off, scalar_mix6_off
      REAL(c_float)  :: c_s, c_k
      INTEGER(c_int) :: mix_full_fields, isfflx
      REAL(c_float)  :: tke_drag_coefficient, tke_heat_flux
      INTEGER(c_int) :: mix_isotropic
      REAL(c_float)  :: mix_upper_bound
      INTEGER(c_int) :: tke_adv_opt
      REAL(c_float)  :: tke_upper_bound
      INTEGER(c_int) :: use_adaptive_time_step
      INTEGER(c_int) :: spec_bdy_width, spec_zone, relax_zone, have_bcs_moist, have_bcs_scalar
   END TYPE wrfgpu_config

   INTERFACE
      INTEGER(c_int) FUNCTION wrfgpu_create(d, c, device, handle) BIND(C, name='wrfgpu_create')
         IMPORT :: c_int, c_ptr, wrfgpu_dims, wrfgpu_config
         TYPE(wrfgpu_dims),   INTENT(IN)  :: d
         TYPE(wrfgpu_config), INTENT(IN)  :: c
         INTEGER(c_int), VALUE            :: device
         TYPE(c_ptr),         INTENT(OUT) :: handle
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_destroy(handle) BIND(C, name='wrfgpu_destroy')
         IMPORT :: c_int, c_ptr
         TYPE(c_ptr), VALUE :: handle
      END FUNCTION
      TYPE(c_ptr) FUNCTION wrfgpu_last_error() BIND(C, name='wrfgpu_last_error')
         IMPORT :: c_ptr
      END FUNCTION
      INTEGER(c_size_t) FUNCTION wrfgpu_field_size(handle, name) BIND(C, name='wrfgpu_field_size')
         IMPORT :: c_size_t, c_ptr, c_char
         TYPE(c_ptr), VALUE :: handle
         CHARACTER(kind=c_char), INTENT(IN) :: name(*)
      END FUNCTION
      ! names: array of C strings (c_loc of NUL-terminated Fortran strings); host: array of c_loc(grid%field)
      INTEGER(c_int) FUNCTION wrfgpu_upload(handle, nfields, names, host) BIND(C, name='wrfgpu_upload')
         IMPORT :: c_int, c_ptr
         TYPE(c_ptr), VALUE :: handle
         INTEGER(c_int), VALUE :: nfields
         TYPE(c_ptr), INTENT(IN) :: names(*), host(*)
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_download(handle, nfields, names, host) BIND(C, name='wrfgpu_download')
         IMPORT :: c_int, c_ptr
         TYPE(c_ptr), VALUE :: handle
         INTEGER(c_int), VALUE :: nfields
         TYPE(c_ptr), INTENT(IN) :: names(*), host(*)
      END FUNCTION
      ! history / restart alarms (share/mediation_integrate.F med_before_solve_io): stage the stream's fields and go on
      ! stepping; wait + wrfgpu_snapshot_host give the pinned host copy that med_hist_out / med_restart_out unpack
      INTEGER(c_int) FUNCTION wrfgpu_snapshot_begin(handle, nfields, names) BIND(C, name='wrfgpu_snapshot_begin')
         IMPORT :: c_int, c_ptr
         TYPE(c_ptr), VALUE :: handle
         INTEGER(c_int), VALUE :: nfields
         TYPE(c_ptr), INTENT(IN) :: names(*)
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_snapshot_wait(handle) BIND(C, name='wrfgpu_snapshot_wait')
         IMPORT :: c_int, c_ptr
         TYPE(c_ptr), VALUE :: handle
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_snapshot_busy(handle) BIND(C, name='wrfgpu_snapshot_busy')
         IMPORT :: c_int, c_ptr
         TYPE(c_ptr), VALUE :: handle
      END FUNCTION
      TYPE(c_ptr) FUNCTION wrfgpu_snapshot_host(handle, name, n) BIND(C, name='wrfgpu_snapshot_host')
         IMPORT :: c_ptr, c_char, c_size_t
         TYPE(c_ptr), VALUE :: handle
         CHARACTER(KIND=c_char), INTENT(IN) :: name(*)
         INTEGER(c_size_t), INTENT(OUT) :: n
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_start_domain(handle) BIND(C, name='wrfgpu_start_domain')
         IMPORT :: c_int, c_ptr
         TYPE(c_ptr), VALUE :: handle
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_step(handle, nsteps) BIND(C, name='wrfgpu_step')
         IMPORT :: c_int, c_ptr
         TYPE(c_ptr), VALUE :: handle
         INTEGER(c_int), VALUE :: nsteps
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_solve_em(handle, n_in, in_names, in_host, n_out, out_names, out_host) &
            BIND(C, name='wrfgpu_solve_em')
         IMPORT :: c_int, c_ptr
         TYPE(c_ptr), VALUE :: handle
         INTEGER(c_int), VALUE :: n_in, n_out
         TYPE(c_ptr), INTENT(IN) :: in_names(*), in_host(*), out_names(*), out_host(*)
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_device_count() BIND(C, name='wrfgpu_device_count')
         IMPORT :: c_int
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_max_cfl(handle, max_vert_cfl, max_horiz_cfl, reset) BIND(C, name='wrfgpu_max_cfl')
         IMPORT :: c_int, c_ptr, c_float
         TYPE(c_ptr), VALUE :: handle
         REAL(c_float), INTENT(OUT) :: max_vert_cfl, max_horiz_cfl
         INTEGER(c_int), VALUE :: reset
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_set_dtbc(handle, dtbc) BIND(C, name='wrfgpu_set_dtbc')
         IMPORT :: c_int, c_ptr, c_float
         TYPE(c_ptr), VALUE :: handle
         REAL(c_float), VALUE :: dtbc
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_sync(handle) BIND(C, name='wrfgpu_sync')
         IMPORT :: c_int, c_ptr
         TYPE(c_ptr), VALUE :: handle
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_comm_unique_id(id128) BIND(C, name='wrfgpu_comm_unique_id')
         IMPORT :: c_int, c_char
         CHARACTER(kind=c_char), INTENT(OUT) :: id128(128)
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_comm_init(handle, id128, rank, nranks) BIND(C, name='wrfgpu_comm_init')
         IMPORT :: c_int, c_ptr, c_char
         TYPE(c_ptr), VALUE :: handle
         CHARACTER(kind=c_char), INTENT(IN) :: id128(128)
         INTEGER(c_int), VALUE :: rank, nranks
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_compute_mesh(ntasks, nproc_x, nproc_y, ntasks_x, ntasks_y) BIND(C, name='wrfgpu_compute_mesh')
         IMPORT :: c_int
         INTEGER(c_int), VALUE :: ntasks, nproc_x, nproc_y
         INTEGER(c_int), INTENT(OUT) :: ntasks_x, ntasks_y
      END FUNCTION
      INTEGER(c_int) FUNCTION wrfgpu_patch_dims(ids, ide, jds, jde, kds, kde, ntasks_x, ntasks_y, mytask, d) &
            BIND(C, name='wrfgpu_patch_dims')
         IMPORT :: c_int, wrfgpu_dims
         INTEGER(c_int), VALUE :: ids, ide, jds, jde, kds, kde, ntasks_x, ntasks_y, mytask
         TYPE(wrfgpu_dims), INTENT(OUT) :: d
      END FUNCTION
   END INTERFACE

CONTAINS

   ! non-zero status -> wrf_error_fatal with the library's message (frame/module_wrf_error.F),
   ! the reference's error convention for this layer (SURVEY 8b)
   SUBROUTINE wrfgpu_check(rc, where)
      INTEGER(c_int), INTENT(IN) :: rc
      CHARACTER(len=*), INTENT(IN) :: where
      CHARACTER(kind=c_char), POINTER :: cmsg(:)
      CHARACTER(len=512) :: msg
      TYPE(c_ptr) :: p
      INTEGER :: n
      IF (rc == 0) RETURN
      msg = ' '
      p = wrfgpu_last_error()
      IF (c_associated(p)) THEN
         CALL c_f_pointer(p, cmsg, [512])
         n = 1
         DO WHILE (n <= 511 .AND. cmsg(n) /= c_null_char)
            msg(n:n) = cmsg(n)
            n = n + 1
         END DO
      END IF
      CALL wrf_error_fatal(TRIM(where)//': '//TRIM(msg))
   END SUBROUTINE wrfgpu_check

END MODULE module_wrfgpu
