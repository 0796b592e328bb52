!+---------------------------------------------------------------------+
!| astr_gpu_mod -- ISO_C_BINDING interface of libastr_gpu.so            |
!| (include/astr_gpu.h).  This is the module a maintainer adds to       |
!| src/ to put the B200 engine behind the existing time loop            |
!| (src/mainloop.F90 time_integration_rk).  It could not be compiled in |
!| the build image (no Fortran compiler): it is delivered as source and |
!| the same symbols are exercised through ctypes (astr_b200/lib.py).    |
!+---------------------------------------------------------------------+
module astr_gpu_mod
  !
  use, intrinsic :: iso_c_binding
  !
  implicit none
  !
  integer(c_int), parameter :: astr_gpu_abi_version=3
  !
  ! struct astr_cfg of include/astr_gpu.h -- same order, same types
  type, bind(c) :: astr_cfg
    integer(c_int) :: abi_version, device
    integer(c_int) :: im,jm,km
    integer(c_int) :: ia,ja,ka
    integer(c_int) :: hm,numq,ndims
    integer(c_int) :: npdc(3)
    integer(c_int) :: is,ie,js,je,ks,ke
    integer(c_int) :: lhomo(3)
    integer(c_int) :: rank(3)
    integer(c_int) :: size(3)
    integer(c_int) :: nbr(6)
    integer(c_int) :: my_rank
    integer(c_int) :: conschm,difschm,scheme_compact,rkscheme
    integer(c_int) :: lfilter,diffterm,nondimen,flowtype
    integer(c_int) :: recon_schem,conschm_explicit,lchardecomp
    integer(c_int) :: bctype(6)
    integer(c_int) :: legacy_sweep,overlap_visc,xchg_nccl
    integer(c_int) :: xchg_timeout_ms,reserved0
    real(c_double) :: alfa_filter
    real(c_double) :: reynolds,mach,prandtl,gamma,ref_tem
    real(c_double) :: const1,const2,const3,const4,const5,const6,const7
    real(c_double) :: tempconst,tempconst1
    real(c_double) :: deltat
    real(c_double) :: twall(6)
    real(c_double) :: pinf
    real(c_double) :: bfacmpld,shkcrt
    real(c_double) :: uinf,vinf,winf,roinf
  end type astr_cfg
  !
  interface
    integer(c_int) function astr_gpu_init(cfg) bind(c,name='astr_gpu_init')
      import :: c_int, astr_cfg
      type(astr_cfg), intent(in) :: cfg
    end function
    integer(c_int) function astr_gpu_sizeof_cfg() bind(c,name='astr_gpu_sizeof_cfg')
      import :: c_int
    end function
    integer(c_int) function astr_gpu_finalize() bind(c,name='astr_gpu_finalize')
      import :: c_int
    end function
    type(c_ptr) function astr_gpu_last_error() bind(c,name='astr_gpu_last_error')
      import :: c_ptr
    end function
    integer(c_int) function astr_gpu_synchronize() bind(c,name='astr_gpu_synchronize')
      import :: c_int
    end function
    integer(c_int) function astr_gpu_comm_unique_id(id) bind(c,name='astr_gpu_comm_unique_id')
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id(128)
    end function
    integer(c_int) function astr_gpu_comm_init(id,nranks,rank) bind(c,name='astr_gpu_comm_init')
      import :: c_int, c_char
      character(kind=c_char), intent(in) :: id(128)
      integer(c_int), value :: nranks, rank
    end function
    integer(c_int) function astr_gpu_set_metrics(dxi,jacob) bind(c,name='astr_gpu_set_metrics')
      import :: c_int, c_double
      real(c_double), intent(in) :: dxi(*), jacob(*)
    end function
    integer(c_int) function astr_gpu_gridgeom(x) bind(c,name='astr_gpu_gridgeom')
      import :: c_int, c_double
      real(c_double), intent(in) :: x(*)
    end function
    integer(c_int) function astr_gpu_set_grid(x) bind(c,name='astr_gpu_set_grid')
      import :: c_int, c_double
      real(c_double), intent(in) :: x(*)
    end function
    integer(c_int) function astr_gpu_set_inflow(vel_in,tmp_in,tmp_prof) bind(c,name='astr_gpu_set_inflow')
      import :: c_int, c_double
      real(c_double), intent(in) :: vel_in(*),tmp_in(*),tmp_prof(*)
    end function
    integer(c_int) function astr_gpu_upload_state(q,rho,vel,prs,tmp) bind(c,name='astr_gpu_upload_state')
      import :: c_int, c_double
      real(c_double), intent(in) :: q(*),rho(*),vel(*),prs(*),tmp(*)
    end function
    integer(c_int) function astr_gpu_download_state(q,rho,vel,prs,tmp) bind(c,name='astr_gpu_download_state')
      import :: c_int, c_double
      real(c_double), intent(out) :: q(*),rho(*),vel(*),prs(*),tmp(*)
    end function
    integer(c_int) function astr_gpu_get_field(id,host) bind(c,name='astr_gpu_get_field')
      import :: c_int, c_double
      integer(c_int), value :: id
      real(c_double), intent(out) :: host(*)
    end function
    integer(c_int) function astr_gpu_set_field(id,host) bind(c,name='astr_gpu_set_field')
      import :: c_int, c_double
      integer(c_int), value :: id
      real(c_double), intent(in) :: host(*)
    end function
    integer(c_int) function astr_gpu_filterq() bind(c,name='astr_gpu_filterq')
      import :: c_int
    end function
    integer(c_int) function astr_gpu_boucon() bind(c,name='astr_gpu_boucon')
      import :: c_int
    end function
    integer(c_int) function astr_gpu_qswap() bind(c,name='astr_gpu_qswap')
      import :: c_int
    end function
    integer(c_int) function astr_gpu_gradcal() bind(c,name='astr_gpu_gradcal')
      import :: c_int
    end function
    integer(c_int) function astr_gpu_rhscal() bind(c,name='astr_gpu_rhscal')
      import :: c_int
    end function
    integer(c_int) function astr_gpu_rk_update(rkstep,deltat) bind(c,name='astr_gpu_rk_update')
      import :: c_int, c_double
      integer(c_int), value :: rkstep
      real(c_double), value :: deltat
    end function
    integer(c_int) function astr_gpu_spongefilter() bind(c,name='astr_gpu_spongefilter')
      import :: c_int
    end function
    integer(c_int) function astr_gpu_updatefvar() bind(c,name='astr_gpu_updatefvar')
      import :: c_int
    end function
    integer(c_int) function astr_gpu_set_sponge(face,beg,end,coef) bind(c,name='astr_gpu_set_sponge')
      import :: c_int, c_double
      integer(c_int), value :: face,beg,end
      real(c_double), intent(in) :: coef(*)
    end function
    ! crash control (src/mainloop.F90:709-1198): counts are per rank, por / psum stay with the caller
    integer(c_int) function astr_gpu_crashcheck(nbad) bind(c,name='astr_gpu_crashcheck')
      import :: c_int, c_long_long
      integer(c_long_long), intent(out) :: nbad
    end function
    integer(c_int) function astr_gpu_databakup(mode,slot,recover_counter) bind(c,name='astr_gpu_databakup')
      import :: c_int
      integer(c_int), value :: mode                 ! 0 'backup', 1 'recovery'
      integer(c_int), intent(out) :: slot,recover_counter
    end function
    integer(c_int) function astr_gpu_crinod_expansion(counter) bind(c,name='astr_gpu_crinod_expansion')
      import :: c_int, c_long_long
      integer(c_long_long), intent(out) :: counter
    end function
    integer(c_int) function astr_gpu_crashfix(ig0,jg0,nfixed) bind(c,name='astr_gpu_crashfix')
      import :: c_int, c_long_long
      integer(c_int), value :: ig0,jg0
      integer(c_long_long), intent(out) :: nfixed
    end function
    ! checkpoint staging: dense node arrays (0:im,0:jm,0:km) of the datasets ro,u1,u2,u3,p,t (readwrite.F90:1974-1984)
    integer(c_int) function astr_gpu_stage_checkpoint(ro,u1,u2,u3,p,t) bind(c,name='astr_gpu_stage_checkpoint')
      import :: c_int, c_double
      real(c_double), intent(out) :: ro(*),u1(*),u2(*),u3(*),p(*),t(*)
    end function
    integer(c_int) function astr_gpu_restore_checkpoint(ro,u1,u2,u3,p,t) bind(c,name='astr_gpu_restore_checkpoint')
      import :: c_int, c_double
      real(c_double), intent(in) :: ro(*),u1(*),u2(*),u3(*),p(*),t(*)
    end function
    ! spg_def='circl': coef = c_loc(sponge_damp_coef(is:ie,js:je,ks:ke)), c_null_ptr when lsponge_loc is false
    integer(c_int) function astr_gpu_set_sponge_global(coef) bind(c,name='astr_gpu_set_sponge_global')
      import :: c_int, c_ptr
      type(c_ptr), value :: coef
    end function
    integer(c_int) function astr_gpu_rk_stage(rkstep,deltat) bind(c,name='astr_gpu_rk_stage')
      import :: c_int, c_double
      integer(c_int), value :: rkstep
      real(c_double), value :: deltat
    end function
    integer(c_int) function astr_gpu_rk_steps(nsteps,deltat) bind(c,name='astr_gpu_rk_steps')
      import :: c_int, c_double
      integer(c_int), value :: nsteps
      real(c_double), value :: deltat
    end function
    integer(c_int) function astr_gpu_dataswap(id,direction) bind(c,name='astr_gpu_dataswap')
      import :: c_int
      integer(c_int), value :: id, direction
    end function
    integer(c_int) function astr_gpu_set_force(force) bind(c,name='astr_gpu_set_force')
      import :: c_int, c_double
      real(c_double), intent(in) :: force(3)
    end function
    integer(c_int) function astr_gpu_reduce_tgv(out) bind(c,name='astr_gpu_reduce_tgv')
      import :: c_int, c_double
      real(c_double), intent(out) :: out(3)
    end function
    ! cflcal (src/commcal.F90:27): block maxima deltai,deltaj,deltak; the caller applies pmax
    integer(c_int) function astr_gpu_reduce_cfl(out) bind(c,name='astr_gpu_reduce_cfl')
      import :: c_int, c_double
      real(c_double), intent(out) :: out(3)
    end function
    ! massfluxchan / fbcxchan (src/statistic.F90:1437,1303): block sums; the caller applies psum and /norm
    integer(c_int) function astr_gpu_reduce_channel(out) bind(c,name='astr_gpu_reduce_channel')
      import :: c_int, c_double
      real(c_double), intent(out) :: out(2)
    end function
  end interface
  !
  contains
  !
  !+-------------------------------------------------------------------+
  !| status check in the reference's error style: print + mpistop      |
  !| (src/parallel.F90:1278).                                          |
  !+-------------------------------------------------------------------+
  subroutine gpu_check(ierr,where)
    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    integer :: n
    if(ierr/=0) then
      call c_f_pointer(astr_gpu_last_error(),msg,[512])
      n=1
      do while(n<512 .and. msg(n)/=c_null_char)
        n=n+1
      enddo
      print*,' !! astr_gpu error @ ',where,': ',msg(1:n-1)
      stop ' !! astr_gpu !!'   ! the shim inside src/ calls mpistop here
    endif
  end subroutine gpu_check
  !
  !+-------------------------------------------------------------------+
  !| replaces the body of solvrinit (src/comsolver.F90:47): fills the   |
  !| cfg from commvar/parallel module data and creates the device side. |
  !+-------------------------------------------------------------------+
  subroutine gpu_solvrinit(im,jm,km,ia,ja,ka,npdci,npdcj,npdck,is,ie,js,je,ks,ke,   &
                           lihomo,ljhomo,lkhomo,irk,jrk,krk,isize,jsize,ksize,     &
                           nbr,mpirank,lfilter,diffterm,alfa_filter,reynolds,mach, &
                           prandtl,gamma,ref_tem,const,tempconst,tempconst1,deltat, &
                           device,flowtype,conschm,difschm,bctype,twall, &
                           recon_schem,lchardecomp,bfacmpld,shkcrt,pinf,ndims,nondimen, &
                           uinf,vinf,winf,roinf,rkscheme)
    character(len=*), intent(in) :: flowtype        ! commvar flowtype: 'channel' enables src_chan
    character(len=3), intent(in), optional :: rkscheme  ! commvar rkscheme: 'rk3' (default) or 'rk4'
    character(len=4), intent(in) :: conschm,difschm ! '643c' or '642e' (comsolver.F90:76-84)
    integer, intent(in) :: bctype(6)                ! commvar bctype(1:6)
    integer, intent(in) :: recon_schem              ! commvar recon_schem
    logical, intent(in) :: lchardecomp              ! commvar lchardecomp
    real(8), intent(in) :: bfacmpld,shkcrt          ! commvar bfacmpld, shkcrt
    real(8), intent(in) :: pinf                     ! commvar pinf
    real(8), intent(in) :: uinf,vinf,winf,roinf     ! commvar free stream (farfield faces)
    integer, intent(in) :: ndims                    ! commvar ndims (3, or 2 with km=0)
    logical, intent(in) :: nondimen                 ! commvar nondimen
    real(8), intent(in) :: twall(6)                 ! commvar twall(1:6)
    integer, intent(in) :: im,jm,km,ia,ja,ka,npdci,npdcj,npdck,is,ie,js,je,ks,ke
    logical, intent(in) :: lihomo,ljhomo,lkhomo,lfilter,diffterm
    integer, intent(in) :: irk,jrk,krk,isize,jsize,ksize,nbr(6),mpirank,device
    real(8), intent(in) :: alfa_filter,reynolds,mach,prandtl,gamma,ref_tem,const(7), &
                           tempconst,tempconst1,deltat
    type(astr_cfg) :: cfg
    if(astr_gpu_sizeof_cfg()/=int(c_sizeof(cfg),c_int)) stop ' !! astr_cfg layout mismatch'
    cfg%abi_version=astr_gpu_abi_version; cfg%device=device
    cfg%im=im; cfg%jm=jm; cfg%km=km; cfg%ia=ia; cfg%ja=ja; cfg%ka=ka
    cfg%hm=5; cfg%numq=5; cfg%ndims=ndims
    cfg%npdc=[npdci,npdcj,npdck]
    cfg%is=is; cfg%ie=ie; cfg%js=js; cfg%je=je; cfg%ks=ks; cfg%ke=ke
    cfg%lhomo=merge(1,0,[lihomo,ljhomo,lkhomo])
    cfg%rank=[irk,jrk,krk]; cfg%size=[isize,jsize,ksize]
    cfg%nbr=nbr          ! mpileft,mpiright,mpidown,mpiup,mpiback,mpifront ; MPI_PROC_NULL -> -1
    cfg%my_rank=mpirank
    read(conschm(1:3),*) cfg%conschm
    read(difschm(1:3),*) cfg%difschm
    cfg%scheme_compact=merge(1,0,difschm(4:4)=='c'); cfg%rkscheme=3
    if(present(rkscheme)) cfg%rkscheme=merge(4,3,rkscheme=='rk4')
    cfg%lfilter=merge(1,0,lfilter); cfg%diffterm=merge(1,0,diffterm)
    cfg%nondimen=merge(1,0,nondimen); cfg%flowtype=merge(1,0,trim(flowtype)=='channel')
    cfg%bctype=bctype; cfg%twall=twall
    cfg%legacy_sweep=0; cfg%overlap_visc=0; cfg%xchg_nccl=0   ! engine switches: defaults
    cfg%xchg_timeout_ms=0; cfg%reserved0=0
    cfg%recon_schem=recon_schem; cfg%lchardecomp=merge(1,0,lchardecomp)
    cfg%conschm_explicit=merge(1,0,conschm(4:4)=='e' .and. mod(cfg%conschm/100,2)==1)
    cfg%bfacmpld=bfacmpld; cfg%shkcrt=shkcrt; cfg%pinf=pinf
    cfg%uinf=uinf; cfg%vinf=vinf; cfg%winf=winf; cfg%roinf=roinf
    cfg%alfa_filter=alfa_filter
    cfg%reynolds=reynolds; cfg%mach=mach; cfg%prandtl=prandtl; cfg%gamma=gamma; cfg%ref_tem=ref_tem
    cfg%const1=const(1); cfg%const2=const(2); cfg%const3=const(3); cfg%const4=const(4)
    cfg%const5=const(5); cfg%const6=const(6); cfg%const7=const(7)
    cfg%tempconst=tempconst; cfg%tempconst1=tempconst1
    cfg%deltat=deltat
    call gpu_check(astr_gpu_init(cfg),'solvrinit')
  end subroutine gpu_solvrinit
  !
end module astr_gpu_mod
