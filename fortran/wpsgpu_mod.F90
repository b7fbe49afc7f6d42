!-----------------------------------------------------------------------------------------------
! wpsgpu_mod -- ISO_C_BINDING interfaces to libwpsgpu.so (include/wpsgpu.h) for the WPS host code.
! Add this file to metgrid/src and geogrid/src, link with -lwpsgpu -lcudart.  (Not compiled in this
! repository's image: it has no Fortran compiler.  See INTEGRATION.md.)
!-----------------------------------------------------------------------------------------------
module wpsgpu_mod

   use iso_c_binding
   implicit none

   integer(c_int), parameter :: WPSGPU_MAX_OPS = 16, WPSGPU_MAX_GAUSS = 2048

   ! proj_info as resolved by wpsgpu_map_set (wpsgpu_proj in wpsgpu.h)
   type, bind(C) :: wpsgpu_proj
      integer(c_int32_t) :: code, nlat, nxmin, nxmax, stagger, comp_ll, init, n_gauss
      real(c_float) :: lat1, lon1, lat0, lon0, dx, dy, latinc, loninc, dlon, stdlon, truelat1, truelat2, hemi, cone, &
                       polei, polej, rsw, rebydx, knowni, knownj, re_m, rho0, nc, bigc
      real(c_float) :: gauss_lat(WPSGPU_MAX_GAUSS)
   end type wpsgpu_proj

   type, bind(C) :: wpsgpu_field_opts
      integer(c_int32_t) :: ops(WPSGPU_MAX_OPS), opts(WPSGPU_MAX_OPS)
      real(c_float) :: missing_value, fill_missing, masked
      real(c_float) :: mask_val(3)
      character(kind=c_char) :: mask_rel(4)
   end type wpsgpu_field_opts

   type, bind(C) :: wpsgpu_slab
      type(c_ptr) :: slab
      integer(c_int32_t) :: minx, maxx, miny, maxy, bdr
      type(c_ptr) :: mask(3)
      integer(c_int32_t) :: grid_id, field_stagger, use_landmask, process_bdy_width
      type(c_ptr) :: r_arr, valid_mask, new_pts
      integer(c_int32_t) :: raw_nx = 0, raw_big_endian = 0   ! raw-record ingest: slab = fg_data%slab as read, see wpsgpu.h
   end type wpsgpu_slab

   interface
      integer(c_int) function wpsgpu_create(device_id, handle) bind(C, name='wpsgpu_create')
         import :: c_int, c_ptr
         integer(c_int), value :: device_id
         type(c_ptr) :: handle
      end function
      integer(c_int) function wpsgpu_destroy(handle) bind(C, name='wpsgpu_destroy')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
      end function
      function wpsgpu_last_error() bind(C, name='wpsgpu_last_error')
         import :: c_ptr
         type(c_ptr) :: wpsgpu_last_error
      end function
      integer(c_int) function wpsgpu_push_source_projection(iproj, stand_lon, truelat1, truelat2, dxkm, dykm, dlat, dlon, &
                        known_x, known_y, known_lat, known_lon, cassini, earth_radius, proj) &
                        bind(C, name='wpsgpu_push_source_projection')
         import :: c_int, c_float, c_ptr, wpsgpu_proj
         integer(c_int), value :: iproj
         real(c_float), value :: stand_lon, truelat1, truelat2, dxkm, dykm, dlat, dlon, known_x, known_y, known_lat, known_lon
         type(c_ptr), value :: cassini, earth_radius      ! c_null_ptr = OPTIONAL argument absent
         type(wpsgpu_proj) :: proj
      end function
      integer(c_int) function wpsgpu_parse_interp_option(str, codes, opts, maxn, n, has_gcell, threshold) &
                        bind(C, name='wpsgpu_parse_interp_option')
         import :: c_int, c_char, c_int32_t, c_float
         character(kind=c_char) :: str(*)
         integer(c_int32_t) :: codes(*), opts(*)
         integer(c_int), value :: maxn
         integer(c_int) :: n, has_gcell
         real(c_float) :: threshold
      end function
      integer(c_int) function wpsgpu_metgrid_set_grid(handle, grid_id, xlat, xlon, sm1, em1, sm2, em2) &
                        bind(C, name='wpsgpu_metgrid_set_grid')
         import :: c_int, c_ptr, c_float
         type(c_ptr), value :: handle
         integer(c_int), value :: grid_id, sm1, em1, sm2, em2
         real(c_float) :: xlat(*), xlon(*)
      end function
      integer(c_int) function wpsgpu_metgrid_set_landmask(handle, grid_id, landmask) bind(C, name='wpsgpu_metgrid_set_landmask')
         import :: c_int, c_ptr, c_float
         type(c_ptr), value :: handle
         integer(c_int), value :: grid_id
         real(c_float) :: landmask(*)
      end function
      integer(c_int) function wpsgpu_metgrid_set_domain(handle, sd1, ed1, sd2, ed2) bind(C, name='wpsgpu_metgrid_set_domain')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: sd1, ed1, sd2, ed2
      end function
      integer(c_int) function wpsgpu_metgrid_enqueue_slab(handle, proj, fo, slab) bind(C, name='wpsgpu_metgrid_enqueue_slab')
         import :: c_int, c_ptr, wpsgpu_proj, wpsgpu_field_opts, wpsgpu_slab
         type(c_ptr), value :: handle
         type(wpsgpu_proj) :: proj
         type(wpsgpu_field_opts) :: fo
         type(wpsgpu_slab) :: slab
      end function
      integer(c_int) function wpsgpu_metgrid_enqueue_slabs(handle, proj, fo, slabs, n) bind(C, name='wpsgpu_metgrid_enqueue_slabs')
         import :: c_int, c_int32_t, c_ptr, wpsgpu_proj, wpsgpu_field_opts, wpsgpu_slab
         type(c_ptr), value :: handle
         type(wpsgpu_proj) :: proj
         type(wpsgpu_field_opts) :: fo
         type(wpsgpu_slab) :: slabs(*)
         integer(c_int32_t), value :: n
      end function
      integer(c_int) function wpsgpu_metgrid_enqueue_gcell_slab(handle, proj, dom_proj, fo, slab) &
                        bind(C, name='wpsgpu_metgrid_enqueue_gcell_slab')
         import :: c_int, c_ptr, wpsgpu_proj, wpsgpu_field_opts, wpsgpu_slab
         type(c_ptr), value :: handle
         type(wpsgpu_proj) :: proj, dom_proj
         type(wpsgpu_field_opts) :: fo
         type(wpsgpu_slab) :: slab
      end function
      integer(c_int) function wpsgpu_fill_missing_levels(handle, nlev, levels, r_arr, valid_mask, nx, ny) &
                        bind(C, name='wpsgpu_fill_missing_levels')
         import :: c_int, c_int32_t, c_ptr
         type(c_ptr), value :: handle
         integer(c_int32_t), value :: nlev, nx, ny
         integer(c_int32_t) :: levels(*)
         type(c_ptr) :: r_arr(*), valid_mask(*)       ! c_loc of each level's r_arr / valid_mask%iarray
      end function
      integer(c_int) function wpsgpu_create_level(handle, nx, ny, fill_const, src_r_arr, src_valid_mask, r_arr, valid_mask) &
                        bind(C, name='wpsgpu_create_level')
         import :: c_int, c_int32_t, c_ptr, c_float
         type(c_ptr), value :: handle, src_r_arr, src_valid_mask, r_arr, valid_mask
         integer(c_int32_t), value :: nx, ny
         real(c_float), value :: fill_const
      end function
      integer(c_int) function wpsgpu_metgrid_flush(handle) bind(C, name='wpsgpu_metgrid_flush')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
      end function
      integer(c_int) function wpsgpu_metgrid_flush_async(handle) bind(C, name='wpsgpu_metgrid_flush_async')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
      end function
      integer(c_int) function wpsgpu_synchronize(handle) bind(C, name='wpsgpu_synchronize')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
      end function
      integer(c_int) function wpsgpu_rotate_winds(handle, idir, proj, u, u_mask, v, v_mask, us1, us2, ue1, ue2, &
                        vs1, vs2, ve1, ve2, xlon_u, xlon_v, nlev) bind(C, name='wpsgpu_rotate_winds')
         import :: c_int, c_ptr, c_float, c_int32_t, wpsgpu_proj
         type(c_ptr), value :: handle
         integer(c_int), value :: idir, us1, us2, ue1, ue2, vs1, vs2, ve1, ve2, nlev
         type(wpsgpu_proj) :: proj
         real(c_float) :: u(*), v(*), xlon_u(*), xlon_v(*)
         integer(c_int32_t) :: u_mask(*), v_mask(*)
      end function
      integer(c_int) function wpsgpu_rotate_winds_ll(handle, idir, proj, u, u_mask, v, v_mask, us1, us2, ue1, ue2, &
                        vs1, vs2, ve1, ve2, xlon_u, xlon_v, xlat_u, xlat_v, nlev) bind(C, name='wpsgpu_rotate_winds_ll')
         import :: c_int, c_ptr, c_float, c_int32_t, wpsgpu_proj
         type(c_ptr), value :: handle
         integer(c_int), value :: idir, us1, us2, ue1, ue2, vs1, vs2, ve1, ve2, nlev
         type(wpsgpu_proj) :: proj
         real(c_float) :: u(*), v(*), xlon_u(*), xlon_v(*), xlat_u(*), xlat_v(*)   ! latitudes: PROJ_CASSINI branches
         integer(c_int32_t) :: u_mask(*), v_mask(*)
      end function
      integer(c_int) function wpsgpu_smooth(handle, opt, npass, array, sx, ex, sy, ey, sz, ez) bind(C, name='wpsgpu_smooth')
         import :: c_int, c_ptr, c_float
         type(c_ptr), value :: handle
         integer(c_int), value :: opt, npass, sx, ex, sy, ey, sz, ez
         real(c_float) :: array(*)
      end function
      ! ... the geogrid entry points (wpsgpu_geogrid_*, wpsgpu_fractions_dominant, wpsgpu_xytoll_grid, wpsgpu_map_factor,
      !     wpsgpu_coriolis, wpsgpu_rotang, wpsgpu_dfdx, wpsgpu_dfdy, wpsgpu_lltoxy, wpsgpu_xytoll) follow the same pattern;
      !     their C prototypes are in include/wpsgpu.h.
   end interface

contains

   ! Turn a non-zero status into the reference's log-and-abort convention (module_debug.F:50,197-204)
   subroutine wpsgpu_check(istatus)
      use module_debug
      integer(c_int), intent(in) :: istatus
      character(len=512) :: msg
      character(kind=c_char), pointer :: cmsg(:)
      integer :: i
      if (istatus /= 0) then
         call c_f_pointer(wpsgpu_last_error(), cmsg, (/512/))
         msg = ' '
         do i = 1, 512
            if (cmsg(i) == c_null_char) exit
            msg(i:i) = cmsg(i)
         end do
         call mprintf(.true., ERROR, 'wpsgpu: %s', s1=trim(msg))
      end if
   end subroutine wpsgpu_check

end module wpsgpu_mod
