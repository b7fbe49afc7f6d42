/* Compiled by tests/test_abi_host.py with plain gcc against include/wpsgpu.h and linked with -lwpsgpu: prints
 * sizeof / offsetof of every struct that crosses the C ABI as one JSON object, and calls two host-only entry points
 * through the linker (not dlopen) to show that a C or Fortran host program links against the library as documented. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include "wpsgpu.h"

#define SZ(T) printf("\"sizeof(%s)\": %zu,\n", #T, sizeof(T))
#define OFF(T, f) printf("\"%s.%s\": %zu,\n", #T, #f, offsetof(T, f))

int main(void)
{
    printf("{\n");
    SZ(wpsgpu_proj);
    OFF(wpsgpu_proj, code); OFF(wpsgpu_proj, nlat); OFF(wpsgpu_proj, nxmin); OFF(wpsgpu_proj, nxmax); OFF(wpsgpu_proj, stagger);
    OFF(wpsgpu_proj, comp_ll); OFF(wpsgpu_proj, init); OFF(wpsgpu_proj, n_gauss); OFF(wpsgpu_proj, lat1); OFF(wpsgpu_proj, lon1);
    OFF(wpsgpu_proj, lat0); OFF(wpsgpu_proj, lon0); OFF(wpsgpu_proj, dx); OFF(wpsgpu_proj, dy); OFF(wpsgpu_proj, latinc);
    OFF(wpsgpu_proj, loninc); OFF(wpsgpu_proj, dlon); OFF(wpsgpu_proj, stdlon); OFF(wpsgpu_proj, truelat1); OFF(wpsgpu_proj, truelat2);
    OFF(wpsgpu_proj, hemi); OFF(wpsgpu_proj, cone); OFF(wpsgpu_proj, polei); OFF(wpsgpu_proj, polej); OFF(wpsgpu_proj, rsw);
    OFF(wpsgpu_proj, rebydx); OFF(wpsgpu_proj, knowni); OFF(wpsgpu_proj, knownj); OFF(wpsgpu_proj, re_m); OFF(wpsgpu_proj, rho0);
    OFF(wpsgpu_proj, nc); OFF(wpsgpu_proj, bigc); OFF(wpsgpu_proj, gauss_lat);
    OFF(wpsgpu_proj, ixdim); OFF(wpsgpu_proj, jydim); OFF(wpsgpu_proj, phi); OFF(wpsgpu_proj, lambda);
    SZ(wpsgpu_map_args);
    OFF(wpsgpu_map_args, present); OFF(wpsgpu_map_args, lat1); OFF(wpsgpu_map_args, lon1); OFF(wpsgpu_map_args, lat0);
    OFF(wpsgpu_map_args, lon0); OFF(wpsgpu_map_args, knowni); OFF(wpsgpu_map_args, knownj); OFF(wpsgpu_map_args, dx);
    OFF(wpsgpu_map_args, dy); OFF(wpsgpu_map_args, latinc); OFF(wpsgpu_map_args, loninc); OFF(wpsgpu_map_args, stdlon);
    OFF(wpsgpu_map_args, truelat1); OFF(wpsgpu_map_args, truelat2); OFF(wpsgpu_map_args, r_earth); OFF(wpsgpu_map_args, nlat);
    OFF(wpsgpu_map_args, nxmin); OFF(wpsgpu_map_args, nxmax); OFF(wpsgpu_map_args, stagger);
    OFF(wpsgpu_map_args, ixdim); OFF(wpsgpu_map_args, jydim); OFF(wpsgpu_map_args, phi); OFF(wpsgpu_map_args, lambda);
    SZ(wpsgpu_field_opts);
    OFF(wpsgpu_field_opts, ops); OFF(wpsgpu_field_opts, opts); OFF(wpsgpu_field_opts, missing_value); OFF(wpsgpu_field_opts, fill_missing);
    OFF(wpsgpu_field_opts, masked); OFF(wpsgpu_field_opts, mask_val); OFF(wpsgpu_field_opts, mask_rel);
    SZ(wpsgpu_slab);
    OFF(wpsgpu_slab, slab); OFF(wpsgpu_slab, minx); OFF(wpsgpu_slab, maxx); OFF(wpsgpu_slab, miny); OFF(wpsgpu_slab, maxy);
    OFF(wpsgpu_slab, bdr); OFF(wpsgpu_slab, mask); OFF(wpsgpu_slab, grid_id); OFF(wpsgpu_slab, field_stagger);
    OFF(wpsgpu_slab, use_landmask); OFF(wpsgpu_slab, process_bdy_width); OFF(wpsgpu_slab, r_arr); OFF(wpsgpu_slab, valid_mask);
    OFF(wpsgpu_slab, new_pts); OFF(wpsgpu_slab, raw_nx); OFF(wpsgpu_slab, raw_big_endian);
    SZ(wpsgpu_geo_field);
    OFF(wpsgpu_geo_field, dom_proj); OFF(wpsgpu_geo_field, istagger); OFF(wpsgpu_geo_field, xlat); OFF(wpsgpu_geo_field, xlon);
    OFF(wpsgpu_geo_field, start_i); OFF(wpsgpu_geo_field, end_i); OFF(wpsgpu_geo_field, start_j); OFF(wpsgpu_geo_field, end_j);
    OFF(wpsgpu_geo_field, start_k); OFF(wpsgpu_geo_field, end_k); OFF(wpsgpu_geo_field, landmask); OFF(wpsgpu_geo_field, field);
    SZ(wpsgpu_geo_level);
    OFF(wpsgpu_geo_level, src_proj); OFF(wpsgpu_geo_level, ilevel); OFF(wpsgpu_geo_level, itype); OFF(wpsgpu_geo_level, ops);
    OFF(wpsgpu_geo_level, opts); OFF(wpsgpu_geo_level, msg_val); OFF(wpsgpu_geo_level, msg_fill_val); OFF(wpsgpu_geo_level, mask_val);
    OFF(wpsgpu_geo_level, has_scale); OFF(wpsgpu_geo_level, scale_factor); OFF(wpsgpu_geo_level, sr_x); OFF(wpsgpu_geo_level, sr_y);
    SZ(wpsgpu_mpas_mesh);
    OFF(wpsgpu_mpas_mesh, nCells); OFF(wpsgpu_mpas_mesh, nVertices); OFF(wpsgpu_mpas_mesh, nEdges); OFF(wpsgpu_mpas_mesh, maxEdges);
    OFF(wpsgpu_mpas_mesh, landmask); OFF(wpsgpu_mpas_mesh, nEdgesOnCell); OFF(wpsgpu_mpas_mesh, cellsOnCell); OFF(wpsgpu_mpas_mesh, verticesOnCell);
    OFF(wpsgpu_mpas_mesh, edgesOnCell); OFF(wpsgpu_mpas_mesh, cellsOnVertex); OFF(wpsgpu_mpas_mesh, cellsOnEdge); OFF(wpsgpu_mpas_mesh, latCell);
    OFF(wpsgpu_mpas_mesh, lonCell); OFF(wpsgpu_mpas_mesh, latVertex); OFF(wpsgpu_mpas_mesh, lonVertex); OFF(wpsgpu_mpas_mesh, latEdge);
    OFF(wpsgpu_mpas_mesh, lonEdge);
    /* linked calls: map_set of a Lambert conformal projection (host arithmetic) and the error string of a rejected one */
    wpsgpu_map_args a;
    wpsgpu_proj p;
    memset(&a, 0, sizeof(a));
    a.present = WPSGPU_A_TRUELAT1 | WPSGPU_A_TRUELAT2 | WPSGPU_A_STDLON | WPSGPU_A_LAT1 | WPSGPU_A_LON1 | WPSGPU_A_KNOWNI | WPSGPU_A_KNOWNJ | WPSGPU_A_DX;
    a.truelat1 = 30.0f; a.truelat2 = 60.0f; a.stdlon = -98.0f; a.lat1 = 34.83f; a.lon1 = -81.03f; a.knowni = 50.5f; a.knownj = 50.5f; a.dx = 30000.0f;
    int rc_ok = wpsgpu_map_set(WPSGPU_PROJ_LC, &a, &p);
    a.present &= ~WPSGPU_A_DX;
    int rc_bad = wpsgpu_map_set(WPSGPU_PROJ_LC, &a, &p);
    printf("\"map_set_ok\": %d, \"map_set_without_dx\": %d, \"cone\": %.9g, \"error_text_length\": %zu\n}\n", rc_ok, rc_bad, (double)p.cone,
           strlen(wpsgpu_last_error()));
    return 0;
}
