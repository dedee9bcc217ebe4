!WRF:MEDIATION_LAYER:SOLVER
!
! solve_em_gpu.F90 -- drop-in replacement for dyn_em/solve_em.F: same subroutine name, same
! argument list (grid, config_flags, <dummy_new_args.inc>), so share/solve_interface.F:50-55 and
! everything above it are unchanged.  dyn_em/Makefile lists this object instead of solve_em.o
! (INTEGRATION.md).  The Registry state stays where the Registry put it -- grid%u_2, grid%t_2,
! grid%mu_2, grid%ph_2, grid%tke_2, moist(:,:,:,P_QV), scalar(:,:,:,n) ... -- and is mirrored in HBM;
! between history / restart alarms the device copy is the live one and the host arrays are refreshed
! only when med_after_solve_io is about to read them: the reference's own test for "an alarm rings at
! the end of this step", Is_alarm_tstep on grid%alarms(HISTORY_ALARM) / (restart_alarm)
! (dyn_em/solve_em.F:360-374, frame/module_domain.F:2484).
!
! Specified / nested lateral boundaries: grid%u_bxs ... grid%mu_btye, moist_b*, scalar_b* and grid%fcx / gcx are uploaded
! whenever the host has refilled them (grid%dtbc == 0); relaxation, the specified rows and their end-of-step reset run on
! the device.
!
! What the device does not do is host-side physics: first_rk_step_part1 (radiation, surface layer,
! PBL, cumulus; dyn_em/solve_em.F:807-871) is not called, so a namelist that selects any of them is
! refused with wrf_error_fatal instead of running without their tendencies.  Diffusion (diff_opt,
! km_opt), Kessler microphysics and the boundary fills run on the device.
!
! Not compiled in the development image (no Fortran compiler).  tests/abi_driver.c replays this file's
! call sequence -- wrfgpu_create with a filled wrfgpu_config, the pointer tables of wrfgpu_upload,
! wrfgpu_start_domain, wrfgpu_step, wrfgpu_download -- from C against the same library, so every entry
! point is exercised the way this shim uses it.
!
SUBROUTINE solve_em ( grid , config_flags  &
#include "dummy_new_args.inc"
                    )
   USE module_state_description
   USE module_domain,          ONLY : domain, get_ijk_from_grid, is_alarm_tstep
   USE module_domain_type,     ONLY : history_alarm, restart_alarm
   USE module_configure,       ONLY : grid_config_rec_type
   USE module_driver_constants
#ifdef DM_PARALLEL
   USE module_dm,              ONLY : ntasks_x, ntasks_y, mytask, local_communicator
#endif
   USE module_wrf_error
   USE module_wrfgpu
   USE, INTRINSIC :: iso_c_binding
   IMPLICIT NONE
#ifdef DM_PARALLEL
   INCLUDE 'mpif.h'
#endif

   TYPE(domain) , TARGET          :: grid
   TYPE (grid_config_rec_type) , INTENT(IN) :: config_flags
#include "dummy_new_decl.inc"

   INTEGER, PARAMETER :: MAXF = 384
   TYPE(c_ptr), SAVE :: handle(max_domains) = c_null_ptr
   LOGICAL,     SAVE :: resident(max_domains) = .FALSE.
   TYPE(wrfgpu_dims)   :: d
   TYPE(wrfgpu_config) :: c
   CHARACTER(len=24, kind=c_char), TARGET, SAVE :: nm(MAXF)
   TYPE(c_ptr) :: names(MAXF), host(MAXF)
   INTEGER :: nf, id, ierr, device, ndev, n, m, nsc, my_rank, nx_tasks, ny_tasks
   INTEGER :: ids, ide, jds, jde, kds, kde, ims, ime, jms, jme, kms, kme, ips, ipe, jps, jpe, kps, kpe
   CHARACTER(kind=c_char) :: ncclid(128)
   CHARACTER(len=16) :: sname
   LOGICAL :: host_needs_state
   REAL(c_float) :: cfl_v, cfl_h

   id = grid%id
   CALL get_ijk_from_grid ( grid , ids, ide, jds, jde, kds, kde, ims, ime, jms, jme, kms, kme, &
                                   ips, ipe, jps, jpe, kps, kpe )
#ifdef DM_PARALLEL
   my_rank = mytask; nx_tasks = ntasks_x; ny_tasks = ntasks_y
#else
   my_rank = 0; nx_tasks = 1; ny_tasks = 1
#endif
   nsc = num_scalar - PARAM_FIRST_SCALAR + 1        ! species of the scalar family (slot 1 is the Registry's dummy)

   IF ( .NOT. c_associated(handle(id)) ) THEN
      ! host-side physics would need first_rk_step_part1/2 and their *_tendf arrays: refuse instead of dropping them
      IF ( config_flags%ra_lw_physics /= 0 .OR. config_flags%ra_sw_physics /= 0 .OR. config_flags%sf_sfclay_physics /= 0 .OR. &
           config_flags%sf_surface_physics /= 0 .OR. config_flags%bl_pbl_physics /= 0 .OR. config_flags%cu_physics /= 0 ) &
         CALL wrf_error_fatal( 'solve_em (wrfgpu): radiation / surface / PBL / cumulus physics run on the host '// &
                               '(first_rk_step_part1) and are not wired into this solver' )
      d = wrfgpu_dims(ids, ide, jds, jde, kds, kde, ims, ime, jms, jme, kms, kme, ips, ipe, jps, jpe, kps, kpe)
      c%periodic_x   = MERGE(1,0,config_flags%periodic_x);   c%periodic_y   = MERGE(1,0,config_flags%periodic_y)
      c%symmetric_xs = MERGE(1,0,config_flags%symmetric_xs); c%symmetric_xe = MERGE(1,0,config_flags%symmetric_xe)
      c%symmetric_ys = MERGE(1,0,config_flags%symmetric_ys); c%symmetric_ye = MERGE(1,0,config_flags%symmetric_ye)
      c%open_xs = MERGE(1,0,config_flags%open_xs); c%open_xe = MERGE(1,0,config_flags%open_xe)
      c%open_ys = MERGE(1,0,config_flags%open_ys); c%open_ye = MERGE(1,0,config_flags%open_ye)
      c%specified = MERGE(1,0,config_flags%specified); c%nested = MERGE(1,0,config_flags%nested)
      c%polar = MERGE(1,0,config_flags%polar)
      c%rk_ord = config_flags%rk_ord;  c%time_step_sound = config_flags%time_step_sound
      c%h_mom_adv_order = config_flags%h_mom_adv_order; c%v_mom_adv_order = config_flags%v_mom_adv_order
      c%h_sca_adv_order = config_flags%h_sca_adv_order; c%v_sca_adv_order = config_flags%v_sca_adv_order
      c%moist_adv_opt = config_flags%moist_adv_opt; c%scalar_adv_opt = config_flags%scalar_adv_opt
      c%momentum_adv_opt = config_flags%momentum_adv_opt
      c%non_hydrostatic = MERGE(1,0,config_flags%non_hydrostatic); c%top_lid = MERGE(1,0,config_flags%top_lid)
      c%damp_opt = config_flags%damp_opt; c%w_damping = config_flags%w_damping
      c%pert_coriolis = MERGE(1,0,config_flags%pert_coriolis)
      c%use_theta_m = config_flags%use_theta_m; c%hypsometric_opt = config_flags%hypsometric_opt
      c%phi_adv_z = config_flags%phi_adv_z; c%zadvect_implicit = config_flags%zadvect_implicit
      c%diff_opt = config_flags%diff_opt; c%km_opt = config_flags%km_opt; c%diff_6th_opt = config_flags%diff_6th_opt
      c%mp_physics = config_flags%mp_physics; c%time_step = config_flags%time_step; c%map_proj = config_flags%map_proj
      c%dt = grid%dt; c%dx = config_flags%dx; c%dy = config_flags%dy
      c%epssm = config_flags%epssm; c%smdiv = config_flags%smdiv; c%emdiv = config_flags%emdiv
      c%dampcoef = config_flags%dampcoef; c%zdamp = config_flags%zdamp; c%w_crit_cfl = config_flags%w_crit_cfl
      c%p_top = grid%p_top; c%cf1 = grid%cf1; c%cf2 = grid%cf2; c%cf3 = grid%cf3; c%cfn = grid%cfn; c%cfn1 = grid%cfn1
      c%khdif = config_flags%khdif; c%kvdif = config_flags%kvdif
      c%num_moist = num_moist - PARAM_FIRST_SCALAR + 1;  c%num_scalar = nsc
      c%ntasks_x = nx_tasks; c%ntasks_y = ny_tasks; c%mytask = my_rank
      c%diff_6th_factor = config_flags%diff_6th_factor; c%diff_6th_thresh = config_flags%diff_6th_thresh
      c%diff_6th_slopeopt = config_flags%diff_6th_slopeopt
      c%moist_mix6_off = MERGE(1,0,config_flags%moist_mix6_off); c%scalar_mix6_off = MERGE(1,0,config_flags%scalar_mix6_off)
      c%c_s = config_flags%c_s; c%c_k = config_flags%c_k; c%mix_full_fields = MERGE(1,0,config_flags%mix_full_fields)
      c%isfflx = config_flags%isfflx; c%tke_drag_coefficient = config_flags%tke_drag_coefficient; c%tke_heat_flux = config_flags%tke_heat_flux
      c%mix_isotropic = config_flags%mix_isotropic; c%mix_upper_bound = config_flags%mix_upper_bound
      c%tke_adv_opt = config_flags%tke_adv_opt; c%tke_upper_bound = config_flags%tke_upper_bound
      c%use_adaptive_time_step = MERGE(1,0,config_flags%use_adaptive_time_step)
      c%spec_bdy_width = config_flags%spec_bdy_width; c%spec_zone = grid%spec_zone; c%relax_zone = grid%relax_zone
      c%have_bcs_moist = MERGE(1,0,config_flags%have_bcs_moist); c%have_bcs_scalar = MERGE(1,0,config_flags%have_bcs_scalar)
      ndev = wrfgpu_device_count()
      IF ( ndev < 1 ) CALL wrf_error_fatal( 'solve_em (wrfgpu): no CUDA device on this node (there is no CPU fallback)' )
      device = MOD(my_rank, ndev)                    ! one rank per GPU of the node, ranks placed node by node
      CALL wrfgpu_check( wrfgpu_create(d, c, device, handle(id)), 'wrfgpu_create' )
#ifdef DM_PARALLEL
      IF ( nx_tasks*ny_tasks > 1 ) THEN             ! bootstrap over the communicator RSL_LITE already has
         IF ( my_rank == 0 ) CALL wrfgpu_check( wrfgpu_comm_unique_id(ncclid), 'wrfgpu_comm_unique_id' )
         CALL MPI_Bcast( ncclid, 128, MPI_BYTE, 0, local_communicator, ierr )
         CALL wrfgpu_check( wrfgpu_comm_init(handle(id), ncclid, my_rank, nx_tasks*ny_tasks), 'wrfgpu_comm_init' )
      END IF
#endif
   END IF

   ! ---- host -> device: everything the step reads, once; afterwards only what the host changed
   IF ( .NOT. resident(id) ) THEN
      nf = 0
      CALL add('u_1',grid%u_1); CALL add('u_2',grid%u_2); CALL add('v_1',grid%v_1); CALL add('v_2',grid%v_2)
      CALL add('w_1',grid%w_1); CALL add('w_2',grid%w_2); CALL add('t_1',grid%t_1); CALL add('t_2',grid%t_2)
      CALL add('ph_1',grid%ph_1); CALL add('ph_2',grid%ph_2); CALL add('phb',grid%phb); CALL add('pb',grid%pb)
      CALL add('alb',grid%alb); CALL add('p',grid%p); CALL add('al',grid%al); CALL add('alt',grid%alt)
      CALL add('h_diabatic',grid%h_diabatic)
      CALL add('t_init',grid%t_init)      ! base state of the theta diffusion (diff_opt = 1)
      CALL add('moist_qv', moist(:,:,:,P_QV))
      IF ( config_flags%mp_physics == 1 ) THEN
         CALL add('moist_qc', moist(:,:,:,P_QC)); CALL add('moist_qr', moist(:,:,:,P_QR))
      END IF
      DO n = 1, nsc                                  ! scalar(:,:,:,PARAM_FIRST_SCALAR+n-1) <-> "scalar_n"
         WRITE(sname,'(A,I0)') 'scalar_', n
         CALL add(TRIM(sname), scalar(:,:,:,PARAM_FIRST_SCALAR+n-1))
      END DO
      IF ( config_flags%km_opt == 2 .AND. config_flags%diff_opt == 2 ) THEN   ! the prognostic tke of the 1.5-order closure
         CALL add('tke_1',grid%tke_1); CALL add('tke_2',grid%tke_2)
      END IF
      CALL add2('mu_1',grid%mu_1); CALL add2('mu_2',grid%mu_2); CALL add2('mub',grid%mub); CALL add2('ht',grid%ht)
      CALL add2('msftx',grid%msftx); CALL add2('msfty',grid%msfty); CALL add2('msfux',grid%msfux); CALL add2('msfuy',grid%msfuy)
      CALL add2('msfvx',grid%msfvx); CALL add2('msfvx_inv',grid%msfvx_inv); CALL add2('msfvy',grid%msfvy)
      CALL add2('f',grid%f); CALL add2('e',grid%e); CALL add2('sina',grid%sina); CALL add2('cosa',grid%cosa)
      IF ( config_flags%mp_physics == 1 ) THEN
         CALL add2('rainnc',grid%rainnc); CALL add2('rainncv',grid%rainncv)
      END IF
      CALL add1('znu',grid%znu); CALL add1('znw',grid%znw); CALL add1('fnm',grid%fnm); CALL add1('fnp',grid%fnp)
      CALL add1('rdnw',grid%rdnw); CALL add1('rdn',grid%rdn); CALL add1('dnw',grid%dnw); CALL add1('dn',grid%dn)
      CALL add1('c1h',grid%c1h); CALL add1('c2h',grid%c2h); CALL add1('c1f',grid%c1f); CALL add1('c2f',grid%c2f)
      CALL add1('c3h',grid%c3h); CALL add1('c4h',grid%c4h); CALL add1('c3f',grid%c3f); CALL add1('c4f',grid%c4f)
      CALL add1('u_base',grid%u_base); CALL add1('v_base',grid%v_base); CALL add1('t_base',grid%t_base)
      CALL add1('qv_base',grid%qv_base); CALL add1('z_base',grid%z_base)
      CALL wrfgpu_check( wrfgpu_upload(handle(id), nf, names, host), 'wrfgpu_upload' )
      CALL wrfgpu_check( wrfgpu_start_domain(handle(id)), 'wrfgpu_start_domain' )
      resident(id) = .TRUE.
   END IF

   ! ---- surface fluxes of the physical-space vertical diffusion: allocation-time zeros without a surface-layer scheme
   !      (a scheme would make this shim refuse above), uploaded so that the device sees what grid% holds
   IF ( config_flags%diff_opt == 2 .AND. config_flags%isfflx /= 0 ) THEN
      nf = 0
      CALL add2('ustm',grid%ustm); CALL add2('hfx',grid%hfx); CALL add2('qfx',grid%qfx)
      CALL wrfgpu_check( wrfgpu_upload(handle(id), nf, names, host), 'wrfgpu_upload(ustm,hfx,qfx)' )
   END IF

   ! ---- specified / nested lateral boundaries: the boundary arrays change when share/input_wrf.F reads the next boundary
   !      time or med_nest_force refills them, and both reset grid%dtbc to 0 (input_wrf.F:1834, mediation_force_domain.F:206);
   !      the device keeps its own dtbc, advanced by wrfgpu_step exactly as solve_em.F:377 advances grid%dtbc below
   IF ( config_flags%specified .OR. config_flags%nested ) THEN
      IF ( grid%dtbc == 0. ) THEN
         nf = 0
         CALL add1('fcx',grid%fcx); CALL add1('gcx',grid%gcx)
         CALL addb('u', grid%u_bxs, grid%u_bxe, grid%u_bys, grid%u_bye, grid%u_btxs, grid%u_btxe, grid%u_btys, grid%u_btye)
         CALL addb('v', grid%v_bxs, grid%v_bxe, grid%v_bys, grid%v_bye, grid%v_btxs, grid%v_btxe, grid%v_btys, grid%v_btye)
         CALL addb('w', grid%w_bxs, grid%w_bxe, grid%w_bys, grid%w_bye, grid%w_btxs, grid%w_btxe, grid%w_btys, grid%w_btye)
         CALL addb('ph', grid%ph_bxs, grid%ph_bxe, grid%ph_bys, grid%ph_bye, grid%ph_btxs, grid%ph_btxe, grid%ph_btys, grid%ph_btye)
         CALL addb('t', grid%t_bxs, grid%t_bxe, grid%t_bys, grid%t_bye, grid%t_btxs, grid%t_btxe, grid%t_btys, grid%t_btye)
         CALL addb('mu', grid%mu_bxs, grid%mu_bxe, grid%mu_bys, grid%mu_bye, grid%mu_btxs, grid%mu_btxe, grid%mu_btys, grid%mu_btye)
         CALL addb('moist_qv', moist_bxs(:,:,:,P_QV), moist_bxe(:,:,:,P_QV), moist_bys(:,:,:,P_QV), moist_bye(:,:,:,P_QV), &
                   moist_btxs(:,:,:,P_QV), moist_btxe(:,:,:,P_QV), moist_btys(:,:,:,P_QV), moist_btye(:,:,:,P_QV))
         IF ( config_flags%mp_physics == 1 .AND. ( config_flags%nested .OR. config_flags%have_bcs_moist ) ) THEN
            CALL addb('moist_qc', moist_bxs(:,:,:,P_QC), moist_bxe(:,:,:,P_QC), moist_bys(:,:,:,P_QC), moist_bye(:,:,:,P_QC), &
                      moist_btxs(:,:,:,P_QC), moist_btxe(:,:,:,P_QC), moist_btys(:,:,:,P_QC), moist_btye(:,:,:,P_QC))
            CALL addb('moist_qr', moist_bxs(:,:,:,P_QR), moist_bxe(:,:,:,P_QR), moist_bys(:,:,:,P_QR), moist_bye(:,:,:,P_QR), &
                      moist_btxs(:,:,:,P_QR), moist_btxe(:,:,:,P_QR), moist_btys(:,:,:,P_QR), moist_btye(:,:,:,P_QR))
         END IF
         IF ( config_flags%nested .OR. config_flags%have_bcs_scalar ) THEN
            DO n = 1, nsc
               WRITE(sname,'(A,I0)') 'scalar_', n
               m = PARAM_FIRST_SCALAR+n-1
               CALL addb(TRIM(sname), scalar_bxs(:,:,:,m), scalar_bxe(:,:,:,m), scalar_bys(:,:,:,m), scalar_bye(:,:,:,m), &
                         scalar_btxs(:,:,:,m), scalar_btxe(:,:,:,m), scalar_btys(:,:,:,m), scalar_btye(:,:,:,m))
            END DO
         END IF
         CALL wrfgpu_check( wrfgpu_upload(handle(id), nf, names, host), 'wrfgpu_upload(boundary arrays)' )
      END IF
      CALL wrfgpu_check( wrfgpu_set_dtbc(handle(id), grid%dtbc), 'wrfgpu_set_dtbc' )
   END IF

   ! ---- the reference's bookkeeping of this call (solve_em.F:360-374)
   host_needs_state = Is_alarm_tstep(grid%domain_clock, grid%alarms(restart_alarm)) .OR. &
                      Is_alarm_tstep(grid%domain_clock, grid%alarms(HISTORY_ALARM))
   grid%itimestep = grid%itimestep + 1
   grid%dtbc = grid%dtbc + grid%dt

   ! ---- the step: dyn_em/solve_em.F:578-4848 on the device
   CALL wrfgpu_check( wrfgpu_step(handle(id), 1), 'wrfgpu_step' )

   IF ( config_flags%use_adaptive_time_step ) THEN     ! solve_em.F:914-930: what adapt_timestep reads next call
      CALL wrfgpu_check( wrfgpu_max_cfl(handle(id), cfl_v, cfl_h, 1), 'wrfgpu_max_cfl' )
      grid%max_vert_cfl = MAX(grid%max_vert_cfl, cfl_v); grid%max_horiz_cfl = MAX(grid%max_horiz_cfl, cfl_h)
      grid%max_cfl_val = MAX(grid%max_cfl_val, grid%max_vert_cfl, grid%max_horiz_cfl)
   END IF

   ! ---- device -> host only when the host is about to read the state (history / restart output after this step)
   IF ( host_needs_state ) THEN
      nf = 0
      CALL add('u_1',grid%u_1); CALL add('u_2',grid%u_2); CALL add('v_1',grid%v_1); CALL add('v_2',grid%v_2)
      CALL add('w_1',grid%w_1); CALL add('w_2',grid%w_2); CALL add('t_1',grid%t_1); CALL add('t_2',grid%t_2)
      CALL add('ph_1',grid%ph_1); CALL add('ph_2',grid%ph_2); CALL add('p',grid%p); CALL add('al',grid%al)
      CALL add('alt',grid%alt); CALL add('ww',grid%ww); CALL add('h_diabatic',grid%h_diabatic)
      CALL add('moist_qv', moist(:,:,:,P_QV))
      IF ( config_flags%mp_physics == 1 ) THEN
         CALL add('moist_qc', moist(:,:,:,P_QC)); CALL add('moist_qr', moist(:,:,:,P_QR))
      END IF
      DO n = 1, nsc
         WRITE(sname,'(A,I0)') 'scalar_', n
         CALL add(TRIM(sname), scalar(:,:,:,PARAM_FIRST_SCALAR+n-1))
      END DO
      IF ( config_flags%km_opt == 2 .AND. config_flags%diff_opt == 2 ) THEN
         CALL add('tke_1',grid%tke_1); CALL add('tke_2',grid%tke_2)
      END IF
      CALL add2('mu_1',grid%mu_1); CALL add2('mu_2',grid%mu_2)
      IF ( config_flags%mp_physics == 1 ) THEN
         CALL add2('rainnc',grid%rainnc); CALL add2('rainncv',grid%rainncv)
      END IF
      CALL wrfgpu_check( wrfgpu_download(handle(id), nf, names, host), 'wrfgpu_download' )
   END IF

CONTAINS
   SUBROUTINE add(name, a)
      CHARACTER(len=*), INTENT(IN) :: name
      REAL, TARGET, CONTIGUOUS, INTENT(IN) :: a(:,:,:)          ! (ims:ime,kms:kme,jms:jme)
      nf = nf + 1; nm(nf) = TRIM(name)//c_null_char; names(nf) = c_loc(nm(nf)); host(nf) = c_loc(a)
   END SUBROUTINE
   SUBROUTINE add2(name, a)
      CHARACTER(len=*), INTENT(IN) :: name
      REAL, TARGET, CONTIGUOUS, INTENT(IN) :: a(:,:)
      nf = nf + 1; nm(nf) = TRIM(name)//c_null_char; names(nf) = c_loc(nm(nf)); host(nf) = c_loc(a)
   END SUBROUTINE
   SUBROUTINE addb(f, bxs, bxe, bys, bye, btxs, btxe, btys, btye)   ! the eight boundary arrays of one field
      CHARACTER(len=*), INTENT(IN) :: f
      REAL, TARGET, CONTIGUOUS, INTENT(IN) :: bxs(:,:,:), bxe(:,:,:), bys(:,:,:), bye(:,:,:), btxs(:,:,:), btxe(:,:,:), btys(:,:,:), btye(:,:,:)
      CALL add(f//'_bxs', bxs);   CALL add(f//'_bxe', bxe);   CALL add(f//'_bys', bys);   CALL add(f//'_bye', bye)
      CALL add(f//'_btxs', btxs); CALL add(f//'_btxe', btxe); CALL add(f//'_btys', btys); CALL add(f//'_btye', btye)
   END SUBROUTINE
   SUBROUTINE add1(name, a)
      CHARACTER(len=*), INTENT(IN) :: name
      REAL, TARGET, CONTIGUOUS, INTENT(IN) :: a(:)
      nf = nf + 1; nm(nf) = TRIM(name)//c_null_char; names(nf) = c_loc(nm(nf)); host(nf) = c_loc(a)
   END SUBROUTINE
END SUBROUTINE solve_em
