!-----------------------------------------------------------------------------------------------
! wpsgpu_mod -- ISO_C_BINDING interfaces to libwpsgpu.so (include/wpsgpu.h) for the WPS host code.
! Add this file to metgrid/src and geogrid/src, link with -lwpsgpu -lcudart.  Every entry point of the header has an
! interface here (tests/test_abi_host.py checks the two lists against each other and the derived types against the C
! struct layouts).  Not compiled in this repository's image: it has no Fortran compiler.  See INTEGRATION.md.
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
      integer(c_int32_t) :: ixdim, jydim      ! PROJ_ROTLL
      real(c_float) :: phi, lambda
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

   ! OPTIONAL arguments of map_set (module_map_utils.F:243-283): bit set in `present` = argument passed
   integer(c_int32_t), parameter :: WPSGPU_A_LAT1 = 1, WPSGPU_A_LON1 = 2, WPSGPU_A_LAT0 = 4, WPSGPU_A_LON0 = 8, &
      WPSGPU_A_KNOWNI = 16, WPSGPU_A_KNOWNJ = 32, WPSGPU_A_DX = 64, WPSGPU_A_DY = 128, WPSGPU_A_LATINC = 256, &
      WPSGPU_A_LONINC = 512, WPSGPU_A_STDLON = 1024, WPSGPU_A_TRUELAT1 = 2048, WPSGPU_A_TRUELAT2 = 4096, &
      WPSGPU_A_NLAT = 8192, WPSGPU_A_NXMIN = 131072, WPSGPU_A_NXMAX = 262144, WPSGPU_A_STAGGER = 524288, &
      WPSGPU_A_R_EARTH = 4194304, WPSGPU_A_IXDIM = 32768, WPSGPU_A_JYDIM = 65536, WPSGPU_A_PHI = 1048576, &
      WPSGPU_A_LAMBDA = 2097152
   integer(c_int), parameter :: WPSGPU_PRECISION_EXACT = 0, WPSGPU_PRECISION_TOLERANCE = 1
   integer(c_int), parameter :: WPSGPU_PTR_HOST = 0, WPSGPU_PTR_DEVICE = 1
   integer(c_int), parameter :: WPSGPU_CONTINUOUS = 0, WPSGPU_CATEGORICAL = 1, WPSGPU_SP_CONTINUOUS = 2

   type, bind(C) :: wpsgpu_map_args
      integer(c_int32_t) :: present      ! uint32_t in C: same size and representation
      real(c_float) :: lat1, lon1, lat0, lon0, knowni, knownj, dx, dy, latinc, loninc, stdlon, truelat1, truelat2, r_earth
      integer(c_int32_t) :: nlat, nxmin, nxmax, stagger
      integer(c_int32_t) :: ixdim, jydim      ! PROJ_ROTLL
      real(c_float) :: phi, lambda
   end type wpsgpu_map_args

   ! calc_field's arguments (process_tile_module.F:1208-1260)
   type, bind(C) :: wpsgpu_geo_field
      type(wpsgpu_proj) :: dom_proj                   ! proj_stack(current_nest_number)
      integer(c_int32_t) :: istagger
      type(c_ptr) :: xlat, xlon                       ! c_loc(xlat_array), c_loc(xlon_array)
      integer(c_int32_t) :: start_i, end_i, start_j, end_j, start_k, end_k
      type(c_ptr) :: landmask                         ! c_null_ptr = optional landmask absent
      type(c_ptr) :: field                            ! c_loc(field)
   end type wpsgpu_geo_field

   ! one priority level of a field (:1271-1345)
   type, bind(C) :: wpsgpu_geo_level
      type(wpsgpu_proj) :: src_proj
      integer(c_int32_t) :: ilevel, itype
      integer(c_int32_t) :: ops(WPSGPU_MAX_OPS), opts(WPSGPU_MAX_OPS)
      real(c_float) :: msg_val, msg_fill_val, mask_val
      integer(c_int32_t) :: has_scale
      real(c_float) :: scale_factor
      integer(c_int32_t) :: sr_x, sr_y
   end type wpsgpu_geo_level

   ! mpas_mesh_type (mpas_mesh.F:5-30) as c_loc's of the module's arrays
   type, bind(C) :: wpsgpu_mpas_mesh
      integer(c_int32_t) :: nCells, nVertices, nEdges, maxEdges
      type(c_ptr) :: landmask, nEdgesOnCell
      type(c_ptr) :: cellsOnCell, verticesOnCell, edgesOnCell
      type(c_ptr) :: cellsOnVertex
      type(c_ptr) :: cellsOnEdge
      type(c_ptr) :: latCell, lonCell, latVertex, lonVertex, latEdge, lonEdge
   end type wpsgpu_mpas_mesh

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
      integer(c_int) function wpsgpu_set_pointer_mode(handle, mode) bind(C, name='wpsgpu_set_pointer_mode')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: mode
      end function
      function wpsgpu_stream(handle) bind(C, name='wpsgpu_stream')
         import :: c_ptr
         type(c_ptr), value :: handle
         type(c_ptr) :: wpsgpu_stream
      end function
      integer(c_int64_t) function wpsgpu_launch_count(handle) bind(C, name='wpsgpu_launch_count')
         import :: c_int64_t, c_ptr
         type(c_ptr), value :: handle
      end function
      integer(c_int) function wpsgpu_map_set(proj_code, args, proj) bind(C, name='wpsgpu_map_set')
         import :: c_int, wpsgpu_map_args, wpsgpu_proj
         integer(c_int), value :: proj_code
         type(wpsgpu_map_args) :: args
         type(wpsgpu_proj) :: proj
      end function
      integer(c_int) function wpsgpu_compute_nest_locations(iproj, coarse, n_domains, parent_id, parent_grid_ratio, &
                        i_parent_start, j_parent_start, ixdim, jydim, projs) bind(C, name='wpsgpu_compute_nest_locations')
         import :: c_int, c_int32_t, wpsgpu_map_args, wpsgpu_proj
         integer(c_int), value :: iproj, n_domains
         type(wpsgpu_map_args) :: coarse
         integer(c_int32_t) :: parent_id(*), parent_grid_ratio(*), i_parent_start(*), j_parent_start(*), ixdim(*), jydim(*)
         type(wpsgpu_proj) :: projs(*)                 ! proj_stack(1:n_domains)
      end function
      integer(c_int) function wpsgpu_lltoxy(handle, proj, stagger, lat, lon, n, x, y) bind(C, name='wpsgpu_lltoxy')
         import :: c_int, c_int64_t, c_ptr, c_float, wpsgpu_proj
         type(c_ptr), value :: handle
         type(wpsgpu_proj) :: proj
         integer(c_int), value :: stagger
         integer(c_int64_t), value :: n
         real(c_float) :: lat(*), lon(*), x(*), y(*)
      end function
      integer(c_int) function wpsgpu_xytoll(handle, proj, stagger, x, y, n, lat, lon) bind(C, name='wpsgpu_xytoll')
         import :: c_int, c_int64_t, c_ptr, c_float, wpsgpu_proj
         type(c_ptr), value :: handle
         type(wpsgpu_proj) :: proj
         integer(c_int), value :: stagger
         integer(c_int64_t), value :: n
         real(c_float) :: x(*), y(*), lat(*), lon(*)
      end function
      integer(c_int) function wpsgpu_metgrid_set_precision(handle, mode) bind(C, name='wpsgpu_metgrid_set_precision')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: mode
      end function
      integer(c_int) function wpsgpu_metgrid_copy_bytes(handle, h2d, d2h) bind(C, name='wpsgpu_metgrid_copy_bytes')
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: handle
         integer(c_int64_t) :: h2d, d2h
      end function
      integer(c_int) function wpsgpu_metgrid_last_kernel(handle, ms, points) bind(C, name='wpsgpu_metgrid_last_kernel')
         import :: c_int, c_ptr, c_float, c_double
         type(c_ptr), value :: handle
         real(c_float) :: ms
         real(c_double) :: points
      end function
      integer(c_int) function wpsgpu_metgrid_last_timings(handle, ms7, points7) bind(C, name='wpsgpu_metgrid_last_timings')
         import :: c_int, c_ptr, c_float, c_double
         type(c_ptr), value :: handle
         real(c_float) :: ms7(7)
         real(c_double) :: points7(7)
      end function
      integer(c_int) function wpsgpu_debug_disable_fast(handle, disable) bind(C, name='wpsgpu_debug_disable_fast')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: disable
      end function
      integer(c_int) function wpsgpu_metgrid_gcell_slab(handle, proj, dom_proj, fo, slab) bind(C, name='wpsgpu_metgrid_gcell_slab')
         import :: c_int, c_ptr, wpsgpu_proj, wpsgpu_field_opts, wpsgpu_slab
         type(c_ptr), value :: handle
         type(wpsgpu_proj) :: proj, dom_proj
         type(wpsgpu_field_opts) :: fo
         type(wpsgpu_slab) :: slab
      end function
      ! ---- geogrid: calc_field (process_tile_module.F:1208-1681) ----
      integer(c_int) function wpsgpu_geogrid_field_begin(handle, f) bind(C, name='wpsgpu_geogrid_field_begin')
         import :: c_int, c_ptr, wpsgpu_geo_field
         type(c_ptr), value :: handle
         type(wpsgpu_geo_field) :: f
      end function
      integer(c_int) function wpsgpu_geogrid_level_begin(handle, l) bind(C, name='wpsgpu_geogrid_level_begin')
         import :: c_int, c_ptr, wpsgpu_geo_level
         type(c_ptr), value :: handle
         type(wpsgpu_geo_level) :: l
      end function
      integer(c_int) function wpsgpu_geogrid_add_tile(handle, tile, src_min_x, src_max_x, src_min_y, src_max_y, src_min_z, &
                        src_max_z, bdr) bind(C, name='wpsgpu_geogrid_add_tile')
         import :: c_int, c_ptr, c_float
         type(c_ptr), value :: handle
         real(c_float) :: tile(*)                      ! where_maps_to_tile%... as get_data_tile returns it
         integer(c_int), value :: src_min_x, src_max_x, src_min_y, src_max_y, src_min_z, src_max_z, bdr
      end function
      integer(c_int) function wpsgpu_geogrid_add_tile_raw(handle, bytes, wordsize, is_signed, endian, row_order, src_min_x, &
                        src_max_x, src_min_y, src_max_y, src_min_z, src_max_z, bdr) bind(C, name='wpsgpu_geogrid_add_tile_raw')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle, bytes            ! c_loc of the buffer the tile file was read into
         integer(c_int), value :: wordsize, is_signed, endian, row_order
         integer(c_int), value :: src_min_x, src_max_x, src_min_y, src_max_y, src_min_z, src_max_z, bdr
      end function
      integer(c_int) function wpsgpu_geogrid_decode_tile(handle, bytes, wordsize, is_signed, endian, row_order, nx, ny, nz, &
                        scalefactor, rarray) bind(C, name='wpsgpu_geogrid_decode_tile')
         import :: c_int, c_ptr, c_float
         type(c_ptr), value :: handle, bytes
         integer(c_int), value :: wordsize, is_signed, endian, row_order, nx, ny, nz
         real(c_float), value :: scalefactor
         real(c_float) :: rarray(*)
      end function
      integer(c_int) function wpsgpu_geogrid_level_end(handle) bind(C, name='wpsgpu_geogrid_level_end')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
      end function
      integer(c_int) function wpsgpu_geogrid_field_finish(handle, processed_domain, n_invalid_category) &
                        bind(C, name='wpsgpu_geogrid_field_finish')
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: handle, processed_domain  ! c_loc(processed_domain%iarray) or c_null_ptr
         integer(c_int64_t) :: n_invalid_category
      end function
      integer(c_int) function wpsgpu_fractions_dominant(handle, field, npts, min_cat, max_cat, msg_fill_val, lm_vals, n_lm, &
                        is_water_mask, landmask_out, dominant_out) bind(C, name='wpsgpu_fractions_dominant')
         import :: c_int, c_int32_t, c_int64_t, c_ptr, c_float
         type(c_ptr), value :: handle, landmask_out, dominant_out   ! c_null_ptr = not wanted
         real(c_float) :: field(*)
         integer(c_int64_t), value :: npts
         integer(c_int), value :: min_cat, max_cat, n_lm, is_water_mask
         real(c_float), value :: msg_fill_val
         integer(c_int32_t) :: lm_vals(*)
      end function
      integer(c_int) function wpsgpu_xytoll_grid(handle, dom, stagger, si, sj, ei, ej, sub_x, sub_y, xlat, xlon) &
                        bind(C, name='wpsgpu_xytoll_grid')
         import :: c_int, c_ptr, c_float, wpsgpu_proj
         type(c_ptr), value :: handle
         type(wpsgpu_proj) :: dom
         integer(c_int), value :: stagger, si, sj, ei, ej
         real(c_float), value :: sub_x, sub_y
         real(c_float) :: xlat(*), xlon(*)
      end function
      integer(c_int) function wpsgpu_map_factor(handle, iproj, truelat1, truelat2, pole_lat, pole_lon, stand_lon, xlat, xlon, &
                        npts, mapfac_x, mapfac_y) bind(C, name='wpsgpu_map_factor')
         import :: c_int, c_int64_t, c_ptr, c_float
         type(c_ptr), value :: handle
         integer(c_int), value :: iproj
         real(c_float), value :: truelat1, truelat2, pole_lat, pole_lon, stand_lon
         integer(c_int64_t), value :: npts
         real(c_float) :: xlat(*), xlon(*), mapfac_x(*), mapfac_y(*)
      end function
      integer(c_int) function wpsgpu_coriolis(handle, xlat, npts, f, e) bind(C, name='wpsgpu_coriolis')
         import :: c_int, c_int64_t, c_ptr, c_float
         type(c_ptr), value :: handle
         integer(c_int64_t), value :: npts
         real(c_float) :: xlat(*), f(*), e(*)
      end function
      integer(c_int) function wpsgpu_rotang(handle, iproj, truelat1, truelat2, stand_lon, xlat, xlon, si, sj, ei, ej, cosa, sina) &
                        bind(C, name='wpsgpu_rotang')
         import :: c_int, c_ptr, c_float
         type(c_ptr), value :: handle
         integer(c_int), value :: iproj, si, sj, ei, ej
         real(c_float), value :: truelat1, truelat2, stand_lon
         real(c_float) :: xlat(*), xlon(*), cosa(*), sina(*)
      end function
      integer(c_int) function wpsgpu_dfdx(handle, src, dst, si, sj, sk, ei, ej, ek, sr_x, mapfac, dom_dx) bind(C, name='wpsgpu_dfdx')
         import :: c_int, c_ptr, c_float
         type(c_ptr), value :: handle, mapfac            ! c_null_ptr = no map factor
         integer(c_int), value :: si, sj, sk, ei, ej, ek, sr_x
         real(c_float), value :: dom_dx
         real(c_float) :: src(*), dst(*)
      end function
      integer(c_int) function wpsgpu_dfdy(handle, src, dst, si, sj, sk, ei, ej, ek, sr_y, mapfac, dom_dy) bind(C, name='wpsgpu_dfdy')
         import :: c_int, c_ptr, c_float
         type(c_ptr), value :: handle, mapfac
         integer(c_int), value :: si, sj, sk, ei, ej, ek, sr_y
         real(c_float), value :: dom_dy
         real(c_float) :: src(*), dst(*)
      end function
      integer(c_int) function wpsgpu_metgrid_set_timing(handle, on) bind(C, name='wpsgpu_metgrid_set_timing')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: on
      end function
      integer(c_int) function wpsgpu_smooth_egrid(handle, opt, npass, array, sx, ex, sy, ey, sz, ez, msgval, hflag) &
                        bind(C, name='wpsgpu_smooth_egrid')
         import :: c_int, c_ptr, c_float
         type(c_ptr), value :: handle
         integer(c_int), value :: opt, npass, sx, ex, sy, ey, sz, ez
         real(c_float) :: array(*)
         real(c_float), value :: msgval, hflag
      end function
      integer(c_int) function wpsgpu_rotate_winds_nmm(handle, idir, proj, u, u_mask, v, v_mask, vs1, vs2, ve1, ve2, xlat_v, xlon_v, nlev) &
                        bind(C, name='wpsgpu_rotate_winds_nmm')
         import :: c_int, c_int32_t, c_ptr, c_float, wpsgpu_proj
         type(c_ptr), value :: handle
         integer(c_int), value :: idir, vs1, vs2, ve1, ve2, nlev
         type(wpsgpu_proj) :: proj
         real(c_float) :: u(*), v(*), xlat_v(*), xlon_v(*)
         integer(c_int32_t) :: u_mask(*), v_mask(*)
      end function
      ! ---- MPAS remapping (remapper.F; process_mpas_fields, process_domain_module.F:1457-1962) ----
      integer(c_int) function wpsgpu_mpas_set_mesh(handle, mesh) bind(C, name='wpsgpu_mpas_set_mesh')
         import :: c_int, c_ptr, wpsgpu_mpas_mesh
         type(c_ptr), value :: handle
         type(wpsgpu_mpas_mesh) :: mesh
      end function
      integer(c_int) function wpsgpu_mpas_remap_setup(handle, grid_id, lat_rad, lon_rad, nlon, nlat) bind(C, name='wpsgpu_mpas_remap_setup')
         import :: c_int, c_ptr, c_float
         type(c_ptr), value :: handle
         integer(c_int), value :: grid_id, nlon, nlat
         real(c_float) :: lat_rad(*), lon_rad(*)
      end function
      integer(c_int) function wpsgpu_mpas_get_tables(handle, grid_id, kind, nearest, sources, weights) bind(C, name='wpsgpu_mpas_get_tables')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: grid_id, kind
         type(c_ptr), value :: nearest, sources, weights      ! c_loc of the remap_info arrays, or c_null_ptr
      end function
      integer(c_int) function wpsgpu_mpas_remap_field(handle, grid_id, decomp, xtype, masked, src, nlev, nw, dst) &
                        bind(C, name='wpsgpu_mpas_remap_field')
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: grid_id, decomp, xtype, masked, nlev, nw
         type(c_ptr), value :: src, dst                       ! c_loc(mpas_field % array2r), c_loc(wrf_field % array3r), ...
      end function
      integer(c_int) function wpsgpu_mpas_remap_planes(handle, grid_id, decomp, masked, src, nlev, mode, landmask, fill_missing, planes) &
                        bind(C, name='wpsgpu_mpas_remap_planes')
         import :: c_int, c_ptr, c_float
         type(c_ptr), value :: handle
         integer(c_int), value :: grid_id, decomp, masked, nlev, mode
         real(c_float) :: src(*)
         type(c_ptr), value :: landmask                       ! c_loc(landmask) or c_null_ptr
         real(c_float), value :: fill_missing
         type(c_ptr) :: planes(*)                             ! c_loc(field_to_store % r_arr) of every level
      end function
      integer(c_int) function wpsgpu_mpas_derive_tt(handle, theta, pressure, nlev, npts) bind(C, name='wpsgpu_mpas_derive_tt')
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: handle
         type(c_ptr) :: theta(*), pressure(*)                ! c_loc(theta_field % r_arr), c_loc(pressure_field % r_arr) per level
         integer(c_int), value :: nlev
         integer(c_int64_t), value :: npts
      end function
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
