"""GPU parity for the host-side calc_field driver (wps_b200/geogrid.py): priority-level recursion, the SP_CONTINUOUS
decision and the tile flood-fill order of process_tile_module.F:1208-1681, on top of libwpsgpu.so -- against the oracle
running the same levels and tiles by hand (accum_* + per-point fallback + finish per level)."""
import numpy as np
import pytest

from tests.util import assert_bitexact, conus_like
from tests.test_gpu_geogrid import cat_tile, topo_tile

pytestmark = pytest.mark.gpu


def oracle_calc_field(orc, capi, odom, stag, xlat, xlon, field, levels, tiles_per_level, landmask, fill):
    """The reference flow by hand: highest priority level first; per level accumulate + point fallback per tile in the
    given order, then the level's post-processing (divide / mask / scale) and processed |= level."""
    from wps_b200.geogrid import effective_type
    gy, gx = xlat.shape
    nxw = (gx + 31) // 32
    processed = np.zeros((gy, nxw), np.int32)
    count = np.zeros_like(field)
    ninv = 0
    for (ilevel, itype, tiles) in tiles_per_level:
        lv = levels[ilevel - 1]
        kw = dict(dlat=lv.dx_deg, dlon=lv.dx_deg, known_x=1.0, known_y=1.0, known_lat=lv.known_lat, known_lon=lv.known_lon)
        osrc = orc.push_source_projection(orc.PROJ_LATLON, **kw)
        level = np.zeros_like(processed)
        count[:] = 0.0
        b = lv.tile_bdr
        use_lm = landmask is not None and stag == 1
        for tx, ty in tiles:
            tile = lv.tile_source(tx, ty)
            if tile is None:
                continue
            if itype == capi.CATEGORICAL:
                ninv += orc.accum_categorical(osrc, odom, stag, tile[0], tx - b, ty - b, b, field, -2, -2, 1, processed, level, ilevel,
                                              lv.missing_value)
            elif itype == capi.SP_CONTINUOUS:
                orc.accum_continuous(osrc, odom, stag, tile, tx - b, ty - b, 1, b, field, count, -2, -2, 1, processed, level, lv.missing_value)
            ninv += orc.tile_point_fallback(osrc, itype, tile, tx - b, ty - b, 1, b, xlat, xlon, landmask, use_lm, lv.masked,
                                            lv.interp_option, lv.missing_value, fill, field, count, -2, -2, 1, processed, level)
        orc.calc_field_finish(itype, field, count, landmask, use_lm, lv.masked, fill, lv.scale_factor, -2, -2, 1, processed, level)
    return processed, ninv


def make_levels(capi, kind):
    from wps_b200.geogrid import SourceLevel
    if kind == "landuse":
        # priority 2: a regional 15-arc-second data set that only exists between 38N-41N, 107W-103W (other tiles: none);
        # priority 1: a global 2-arc-minute set.  The overlay level exercises the half-area rule after every tile.
        def regional(tx, ty):
            lon = -179.9979 + (tx - 1) * 0.00416667
            lat = -89.9979 + (ty - 1) * 0.00416667
            if not (-107.5 <= lon <= -103.0 and 37.5 <= lat <= 41.0):
                return None
            return cat_tile(tx, ty, 240, 1)
        return [SourceLevel(0.0333333, -89.98333, -179.98333, 120, 120, 0, capi.CATEGORICAL, "nearest_neighbor",
                            lambda tx, ty: cat_tile(tx, ty, 120, 1, ncat=21), missing_value=255.0),
                SourceLevel(0.00416667, -89.9979, -179.9979, 240, 240, 0, capi.CATEGORICAL, "nearest_neighbor", regional, missing_value=255.0)]
    # topography: 30-arc-second over everything with average_gcell(4.0): at 9 km it becomes SP_CONTINUOUS (4 x 0.93 km <= 9 km);
    # priority 2: a 5-arc-minute set whose average_gcell threshold is NOT met (4 x 9.3 km > 9 km): stays CONTINUOUS (per point)
    return [SourceLevel(0.00833333, -89.99583, -179.99583, 120, 120, 3, capi.CONTINUOUS, "average_gcell(4.0)+four_pt+average_4pt",
                        lambda tx, ty: topo_tile(tx - 3, ty - 3, 126, 1), missing_value=-32768.0),
            SourceLevel(0.0833333, -89.95833, -179.95833, 30, 30, 2, capi.CONTINUOUS, "average_gcell(4.0)+four_pt+average_4pt",
                        lambda tx, ty: (topo_tile(tx - 2, ty - 2, 34, 1) * 0.5 if (tx // 30 + ty // 30) % 3 else None), missing_value=-32768.0)]


@pytest.mark.parametrize("kind", ["landuse", "topography"])
@pytest.mark.parametrize("order", ["reference", "scan"])
def test_calc_field_driver_two_priority_levels(orc, handle, kind, order):
    from wps_b200 import capi
    from wps_b200.geogrid import calc_field, effective_type
    orc.set_transc_mode(1)
    nx, ny, dx = 44, 38, 9000.0
    odom, gdom = conus_like(nx, ny, dx, 30.0, 60.0, -98.0, 39.0, -105.0)
    gx, gy = nx + 6, ny + 6
    xlat, xlon = orc.get_lat_lon_fields(odom, -2, -2, gx, gy, orc.M)
    levels = make_levels(capi, kind)
    nk = 21 if kind == "landuse" else 1
    rng = np.random.default_rng(3)
    init = rng.uniform(-1, 1, (nk, gy, gx)).astype(np.float32)
    landmask = (rng.random((gy, gx)) < 0.8).astype(np.float32) if kind == "topography" else None
    for lv in levels:
        lv.masked = 0.0 if kind == "topography" else -1.0
    fill = -999.0 if kind == "topography" else 1e20
    field_g = init.copy()
    trace = []
    proc_g = np.zeros((gy, (gx + 31) // 32), np.int32)
    ninv_g = calc_field(handle, gdom, capi.M, xlat, xlon, -2, -2, field_g, levels, dx, landmask=landmask, msg_fill_val=fill, order=order,
                        processed_out=proc_g, trace=trace)
    # the driver's decisions
    assert [t[0] for t in trace] == [2, 1]
    if kind == "topography":
        assert effective_type(levels[0], dx) == capi.SP_CONTINUOUS and effective_type(levels[1], dx) == capi.CONTINUOUS
        assert [t[1] for t in trace] == [capi.CONTINUOUS, capi.SP_CONTINUOUS]
    assert all(len(t[2]) >= 2 for t in trace), "each level reaches several tiles"
    if order == "reference":
        # the flood fill starts at the patch's first point: its tile comes first, and every tile appears once
        for _, _, tiles in trace:
            assert len(set(tiles)) == len(tiles)
    field_o = init.copy()
    proc_o, ninv_o = oracle_calc_field(orc, capi, odom, orc.M, xlat, xlon, field_o, levels, trace, landmask, fill)
    assert (proc_g == proc_o).all(), "processed_domain"
    assert ninv_g == ninv_o
    assert_bitexact(field_g, field_o, "calc_field driver, %s, %s tile order" % (kind, order))
    assert (proc_g != 0).any()
