! cube_gpu.f90 -- ISO_C_BINDING interface to libcubegpu.so (include/cube_gpu.h).
!
! Shipped as source: no Fortran compiler exists in the build image, so this file is not compiled
! or tested here (INTEGRATION.md).  It is the module a maintainer adds to CUBE/main so that the
! step loop of cafcube.f90:25-46 calls the B200 library instead of update_particle / buffer_* /
! particle_mesh.  Every interface names the reference routine it replaces.
module cube_gpu
  use iso_c_binding
  implicit none

  ! struct cube_params (include/cube_gpu.h) -- all default-kind C ints/floats, no padding
  type, bind(C) :: cube_params
    integer(c_int32_t) :: nn(3)        ! images per dimension            parameters.f90:20
    integer(c_int32_t) :: rank         ! this_image()-1                  parameters.f90:180
    integer(c_int32_t) :: nnt          ! parameters.f90:22
    integer(c_int32_t) :: nc           ! parameters.f90:23
    integer(c_int32_t) :: ncell        ! parameters.f90:21 (4)
    integer(c_int32_t) :: ncb          ! parameters.f90:46 (6)
    integer(c_int32_t) :: izipx, izipv ! universe*.fh (2,2)
    integer(c_int32_t) :: np_nc        ! parameters.f90:55
    real(c_float)      :: image_buffer ! parameters.f90:59
    real(c_float)      :: tile_buffer  ! parameters.f90:60
    integer(c_int32_t) :: device       ! CUDA device ordinal
    integer(c_int32_t) :: fine_batch   ! 0 = automatic
    integer(c_int32_t) :: local_group  ! 0: one process per image; k>0: images are threads of this process (in-process group k)
    integer(c_int32_t) :: reserved(3)
  end type

  interface
    ! initialize.f90:1-56 (geometry, FFT plans, kernel_f, kernel_c)
    integer(c_int) function cube_gpu_init(p, fk_table, ck_table, tanf_lut, nccl_unique_id, h) bind(C, name="cube_gpu_init")
      import :: c_int, c_ptr, c_float, cube_params
      type(cube_params), intent(in) :: p
      real(c_float), intent(in) :: fk_table(16,16,16,3), ck_table(3,4,4,4), tanf_lut(0:65535)
      type(c_ptr), value :: nccl_unique_id   ! c_null_ptr for one image
      type(c_ptr), intent(out) :: h
    end function
    ! 128-byte NCCL id made on image 1 and broadcast by the coarray side before cube_gpu_init
    integer(c_int) function cube_gpu_nccl_unique_id(id) bind(C, name="cube_gpu_nccl_unique_id")
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id(128)
    end function
    ! particle_initialization.f90:11-72
    integer(c_int) function cube_gpu_upload(h, xp, vp, rhoc_phys, vfield_phys, nplocal, npglobal, sigma_vi) bind(C, name="cube_gpu_upload")
      import :: c_int, c_ptr, c_int32_t, c_int64_t, c_float
      type(c_ptr), value :: h
      type(c_ptr), value :: xp, vp                       ! c_loc(xp), c_loc(vp): integer(izipx) xp(3,*), integer(izipv) vp(3,*)
      integer(c_int32_t), intent(in) :: rhoc_phys(*)     ! rhoc(1:nt,1:nt,1:nt,:,:,:) contiguous copy
      real(c_float), intent(in) :: vfield_phys(*)        ! vfield(:,1:nt,1:nt,1:nt,:,:,:)
      integer(c_int64_t), value :: nplocal, npglobal
      real(c_float), value :: sigma_vi
    end function
    ! the same, returning while xp and vp (page-locked, untouched until cube_gpu_update_x has returned) are still being copied
    integer(c_int) function cube_gpu_upload_begin(h, xp, vp, rhoc_phys, vfield_phys, nplocal, npglobal, sigma_vi) bind(C, name="cube_gpu_upload_begin")
      import :: c_int, c_ptr, c_int32_t, c_int64_t, c_float
      type(c_ptr), value :: h
      type(c_ptr), value :: xp, vp                       ! c_loc(xp), c_loc(vp): integer(izipx) xp(3,*), integer(izipv) vp(3,*)
      integer(c_int32_t), intent(in) :: rhoc_phys(*)     ! rhoc(1:nt,1:nt,1:nt,:,:,:) contiguous copy
      real(c_float), intent(in) :: vfield_phys(*)        ! vfield(:,1:nt,1:nt,1:nt,:,:,:)
      integer(c_int64_t), value :: nplocal, npglobal
      real(c_float), value :: sigma_vi
    end function
    ! update_particle.f90:1-213
    integer(c_int) function cube_gpu_update_x(h, dt_old, dt, nplocal, sigma_vi_new, std_vsim, overhead_tile) bind(C, name="cube_gpu_update_x")
      import :: c_int, c_ptr, c_int64_t, c_float, c_double
      type(c_ptr), value :: h
      real(c_float), value :: dt_old, dt
      integer(c_int64_t), intent(out) :: nplocal
      real(c_float), intent(out) :: sigma_vi_new, overhead_tile
      real(c_double), intent(out) :: std_vsim(3)         ! std_vsim, std_vsim_c, std_vsim_res
    end function
    ! buffer_density.f90 / buffer_x.f90 / buffer_v.f90
    integer(c_int) function cube_gpu_buffer(h, do_density, do_x, do_v, overhead_image) bind(C, name="cube_gpu_buffer")
      import :: c_int, c_ptr, c_float
      type(c_ptr), value :: h
      integer(c_int), value :: do_density, do_x, do_v
      real(c_float), intent(out) :: overhead_image
    end function
    ! pm.f90:1-247
    integer(c_int) function cube_gpu_particle_mesh(h, a_mid, dt, dt_fine, dt_coarse, dt_vmax, vmax) bind(C, name="cube_gpu_particle_mesh")
      import :: c_int, c_ptr, c_float
      type(c_ptr), value :: h
      real(c_float), value :: a_mid, dt
      real(c_float), intent(out) :: dt_fine, dt_coarse, dt_vmax, vmax
    end function
    ! checkpoint.f90:33-70
    integer(c_int) function cube_gpu_download(h, xp, vp, rhoc_phys, vfield_phys, nplocal, sigma_vi) bind(C, name="cube_gpu_download")
      import :: c_int, c_ptr, c_int32_t, c_int64_t, c_float
      type(c_ptr), value :: h
      type(c_ptr), value :: xp, vp                       ! c_loc of integer(izipx) xp(3,*), integer(izipv) vp(3,*); c_null_ptr = skip
      integer(c_int32_t), intent(out) :: rhoc_phys(*)
      real(c_float), intent(out) :: vfield_phys(*)
      integer(c_int64_t), intent(out) :: nplocal
      real(c_float), intent(out) :: sigma_vi
    end function
    ! positions / velocities (and the per-cell arrays) streamed to page-locked host buffers while later calls run:
    ! xp, rhoc and vfield are final once cube_gpu_update_x has returned (checkpoint.f90:35-50); cube_gpu_download with
    ! c_null_ptr for what was streamed waits for the copies
    integer(c_int) function cube_gpu_download_async(h, xp, vp) bind(C, name="cube_gpu_download_async")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, xp, vp
    end function
    integer(c_int) function cube_gpu_download_cells_async(h, rhoc_phys, vfield_phys) bind(C, name="cube_gpu_download_cells_async")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, rhoc_phys, vfield_phys
    end function
    ! registers a page-locked buffer: the next cube_gpu_particle_mesh streams every tile batch's final velocities into it
    integer(c_int) function cube_gpu_stream_vp(h, vp) bind(C, name="cube_gpu_stream_vp")
      import :: c_int, c_ptr
      type(c_ptr), value :: h, vp
    end function
    ! -DPID: IDs of the particles of the last cube_gpu_upload (file order); they follow every cube_gpu_update_x and cross images
    ! with vp in cube_gpu_buffer(do_v) (buffer_v.f90:23,42,62,81,104)
    integer(c_int) function cube_gpu_upload_pid(h, pid) bind(C, name="cube_gpu_upload_pid")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: h
      integer(c_int64_t), intent(in) :: pid(*)
    end function
    integer(c_int) function cube_gpu_download_pid(h, pid) bind(C, name="cube_gpu_download_pid")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: h
      integer(c_int64_t), intent(out) :: pid(*)
    end function
    ! cicpower + powerspectrum (CUBE/utilities/cicpower.f90, powerspectrum.f90, linear_kbin) of the resident state: xi(nbin,10)
    ! in Fortran order = the reference's xi(10,nbin) transposed; nbin = nint(nyquist*sqrt(3.)) comes back in nbin
    integer(c_int) function cube_gpu_power_spectrum(h, box, xi, nbin_cap, nbin) bind(C, name="cube_gpu_power_spectrum")
      import :: c_int, c_ptr, c_float, c_double
      type(c_ptr), value :: h
      real(c_float), value :: box
      real(c_double), intent(out) :: xi(*)
      integer(c_int), value :: nbin_cap
      integer(c_int), intent(out) :: nbin
    end function
    ! CUBEnu's bookkeeping: nlayer colour passes of update_xp (update_particle.f90:37,55-58) and vmax(3) (pm.f90:349,398)
    integer(c_int) function cube_gpu_set_drift_layers(h, nlayer) bind(C, name="cube_gpu_set_drift_layers")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: nlayer
    end function
    integer(c_int) function cube_gpu_get_vmax3(h, vmax3) bind(C, name="cube_gpu_get_vmax3")
      import :: c_int, c_ptr, c_float
      type(c_ptr), value :: h
      real(c_float), intent(out) :: vmax3(3)
    end function
    ! a second species (-DNEUTRINOS): one handle per species; particle_mesh for both on the first one's meshes
    integer(c_int) function cube_gpu_set_mass_p(h, mass_p) bind(C, name="cube_gpu_set_mass_p")
      import :: c_int, c_ptr, c_float
      type(c_ptr), value :: h
      real(c_float), value :: mass_p
    end function
    integer(c_int) function cube_gpu_particle_mesh_species(h, h2, a_mid, dt, dt_fine, dt_coarse, dt_vmax, vmax, dt_vmax2, vmax2) &
        bind(C, name="cube_gpu_particle_mesh_species")
      import :: c_int, c_ptr, c_float
      type(c_ptr), value :: h, h2
      real(c_float), value :: a_mid, dt
      real(c_float), intent(out) :: dt_fine, dt_coarse, dt_vmax, vmax, dt_vmax2, vmax2
    end function
    integer(c_int) function cube_gpu_finalize(h) bind(C, name="cube_gpu_finalize")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    type(c_ptr) function cube_gpu_last_error() bind(C, name="cube_gpu_last_error")
      import :: c_ptr
    end function
  end interface

contains

  ! reference error convention: print and stop (update_particle.f90:61-67)
  subroutine cube_gpu_check(rc)
    integer(c_int), intent(in) :: rc
    character(kind=c_char), pointer :: msg(:)
    integer :: n
    if (rc == 0) return
    call c_f_pointer(cube_gpu_last_error(), msg, [1024])
    n = 1
    do while (n < 1024 .and. msg(n) /= c_null_char)
      n = n + 1
    end do
    print*, msg(1:n-1)
    error stop
  end subroutine

  ! tan((pi*real(v))/real(nvbin-1)) for every pattern of an integer(izipv) code (nvbin = 2**(8*izipv) entries), evaluated by THIS
  ! build's libm, so that the GPU decodes velocities exactly like pm.f90:102 / update_particle.f90:42 would on this host
  subroutine cube_gpu_make_tanf_lut(lut, izipv)
    integer, intent(in) :: izipv
    real(c_float), intent(out) :: lut(0:2**(8*izipv)-1)
    real, parameter :: pi = 4*atan(1.)
    integer :: u, v, nvbin
    nvbin = 2**(8*izipv)
    do u = 0, nvbin-1
      v = u
      if (u >= nvbin/2) v = u - nvbin      ! the pattern read as a signed integer(izipv)
      lut(u) = tan((pi*real(v))/real(nvbin-1))
    end do
  end subroutine

end module cube_gpu
