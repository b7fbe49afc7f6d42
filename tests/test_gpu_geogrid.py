"""GPU parity: geogrid calc_field pieces (accumulators with the quadtree rule, per-point fallback, finish)."""
import numpy as np
import pytest

from tests.util import assert_bitexact, conus_like, ulp_diff

pytestmark = pytest.mark.gpu


def tiles_for(domain_ll, dxdeg, tile_n, bdr, known_lat=-89.99583, known_lon=-179.99583):
    """regular_ll dataset like topo_30s / modis_landuse_30s: global index space, tile_n x tile_n tiles."""
    lat0, lat1, lon0, lon1 = domain_ll
    ix0 = int(np.floor((lon0 - known_lon) / dxdeg)) + 1
    ix1 = int(np.ceil((lon1 - known_lon) / dxdeg)) + 1
    iy0 = int(np.floor((lat0 - known_lat) / dxdeg)) + 1
    iy1 = int(np.ceil((lat1 - known_lat) / dxdeg)) + 1
    tiles = []
    ty = tile_n * ((iy0 - 1) // tile_n) + 1
    while ty <= iy1:
        tx = tile_n * ((ix0 - 1) // tile_n) + 1
        while tx <= ix1:
            tiles.append((tx - bdr, ty - bdr))
            tx += tile_n
        ty += tile_n
    return tiles


def run_field(orc, handle, itype, chain, nk, start_k, tile_fn, tile_n, bdr, dxdeg, nx=60, ny=50, dx=9000.0,
              landmask=None, mask_val=-1.0, msg=1e20, fill=1e20, ilevel=1, scale=None, stag=1, sr=1, dom=None):
    from wps_b200 import capi
    orc.set_transc_mode(1)
    odom, gdom = dom if dom is not None else conus_like(nx, ny, dx, 30.0, 60.0, -98.0, 39.0, -105.0)
    gx, gy = nx + 6, ny + 6  # patch +- HALO_WIDTH as process_tile does
    xlat, xlon = orc.get_lat_lon_fields(odom, -2, -2, gx, gy, stag)
    kw = dict(dlat=dxdeg, dlon=dxdeg, known_x=1.0, known_y=1.0, known_lat=-89.99583, known_lon=-179.99583)
    osrc, gsrc = orc.push_source_projection(orc.PROJ_LATLON, **kw), capi.push_source_projection(capi.PROJ_LATLON, **kw)
    dom_ll = (float(xlat.min()) - 0.3, float(xlat.max()) + 0.3, float(xlon.min()) - 0.3, float(xlon.max()) + 0.3)
    tiles = tiles_for(dom_ll, dxdeg, tile_n, bdr)
    rng = np.random.default_rng(99)
    init = rng.uniform(-1, 1, (nk, gy, gx)).astype(np.float32)  # "uninitialised" field content, same on both sides
    field_g = init.copy()
    nz = 1 if itype == capi.CATEGORICAL else nk
    handle.geogrid_field_begin(gdom, stag, xlat, xlon, -2, -2, start_k, field_g, landmask)
    handle.geogrid_level_begin(gsrc, ilevel, itype, chain, msg, fill, mask_val, scale, sr, sr)
    field_o = init.copy()
    count_o = np.zeros_like(field_o)
    nxw = (gx + 31) // 32
    processed = np.zeros((gy, nxw), np.int32)
    level = np.zeros((gy, nxw), np.int32)
    tn = tile_n + 2 * bdr
    ninv = 0
    for (tx, ty) in tiles:
        tile = tile_fn(tx, ty, tn, nz)
        handle.geogrid_add_tile(tile, tx, ty, start_k if itype != capi.CATEGORICAL else 1, bdr)
        if itype == capi.CATEGORICAL:
            ninv += orc.accum_categorical(osrc, odom, stag, tile[0], tx, ty, bdr, field_o, -2, -2, start_k, processed, level,
                                          ilevel, msg, float(sr), float(sr))
        elif itype == capi.SP_CONTINUOUS:
            orc.accum_continuous(osrc, odom, stag, tile, tx, ty, start_k, bdr, field_o, count_o, -2, -2, start_k, processed,
                                 level, msg, float(sr), float(sr))
        ninv += orc.tile_point_fallback(osrc, itype, tile, tx, ty, start_k if itype != capi.CATEGORICAL else 1, bdr, xlat, xlon,
                                        landmask, landmask is not None and stag == 1, mask_val, chain, msg, fill, field_o,
                                        count_o, -2, -2, start_k, processed, level)
    handle.geogrid_level_end()
    proc_g = np.zeros_like(processed)
    ninv_g = handle.geogrid_field_finish(proc_g)
    orc.calc_field_finish(itype, field_o, count_o, landmask, landmask is not None and stag == 1, mask_val, fill, scale,
                          -2, -2, start_k, processed, level)
    assert (proc_g == processed).all(), "processed_domain bits differ"
    assert ninv_g == ninv
    return field_g, field_o, len(tiles)


def cat_tile(tx, ty, tn, nz, ncat=21, missing=None):
    j, i = np.mgrid[ty:ty + tn, tx:tx + tn]
    patch = ((i // 7) * 31 + (j // 5) * 17) % ncat + 1
    noise = (i * 7919 + j * 104729) % 13 == 0
    t = np.where(noise, (i + j) % ncat + 1, patch).astype(np.float32)
    if missing is not None:
        t[(i * 31 + j * 57) % 11 == 0] = missing
    return t[None]


def topo_tile(tx, ty, tn, nz):
    j, i = np.mgrid[ty:ty + tn, tx:tx + tn]
    h = np.round(1500 + 1200 * np.sin(i / 37.0) * np.cos(j / 53.0) + ((i * 13 + j * 7) % 17))
    return np.stack([h + 10 * k for k in range(nz)]).astype(np.float32)


def test_categorical_landuse(orc, handle):
    """30-arcsec 21-category land use onto a 9-km Lambert patch: counts are integer-exact."""
    from wps_b200 import capi
    g, o, nt = run_field(orc, handle, capi.CATEGORICAL, "nearest_neighbor", 21, 1, cat_tile, 120, 0, 0.00833333)
    assert nt >= 4
    assert_bitexact(g, o, "LANDUSEF counts")


def test_categorical_with_missing_landmask_and_invalid(orc, handle):
    from wps_b200 import capi
    rng = np.random.default_rng(1)
    lm = (rng.random((56, 66)) < 0.7).astype(np.float32)
    fn = lambda tx, ty, tn, nz: cat_tile(tx, ty, tn, nz, ncat=23, missing=np.float32(255.0))
    g, o, _ = run_field(orc, handle, capi.CATEGORICAL, "nearest_neighbor", 21, 1, fn, 150, 0, 0.00833333, landmask=lm,
                        mask_val=0.0, msg=255.0)
    assert_bitexact(g, o, "categorical with missing / invalid categories / landmask")


def test_categorical_coarse_source_point_fallback(orc, handle):
    """Source coarser than the domain (10-arcmin -> 9 km is ~2 pixels per cell... use 0.25 deg): most cells get no
    pixel from accum_categorical and are filled by get_point + one-hot."""
    from wps_b200 import capi
    g, o, _ = run_field(orc, handle, capi.CATEGORICAL, "nearest_neighbor", 21, 1, cat_tile, 40, 0, 0.25, nx=70, ny=60, dx=9000.0)
    assert_bitexact(g, o, "coarse categorical")


def test_categorical_overlay_half_area_rule(orc, handle):
    """priority level > 1: cells covered by less than half a cell of source data are cleared after each tile."""
    from wps_b200 import capi
    g, o, _ = run_field(orc, handle, capi.CATEGORICAL, "nearest_neighbor", 21, 1, cat_tile, 64, 0, 0.05, nx=40, ny=36,
                        dx=9000.0, ilevel=2)
    assert_bitexact(g, o, "overlay categorical")


def test_nmm_egrid_domain(orc, handle):
    """A rotated lat-lon (NMM E-grid, PROJ_ROTLL) domain: where_maps_to goes through llij_rotlatlon with the HH / VV stagger the
    field is defined on (proc_point_module.F:245-310 -> llxy_module.F:800-804); categorical counts on mass points and
    grid-cell averages on velocity points, against the oracle."""
    from wps_b200 import capi
    kw = dict(ixdim=40, jydim=61, phi=2.4, lat1=39.0, lon1=-100.0, stagger=capi.HH)
    kw["lambda"] = 3.0
    orc.set_transc_mode(1)
    dom = (orc.map_set(orc.PROJ_ROTLL, **kw), capi.map_set(capi.PROJ_ROTLL, **kw))
    g, o, nt = run_field(orc, handle, capi.CATEGORICAL, "nearest_neighbor", 21, 1, cat_tile, 120, 0, 0.00833333, nx=39, ny=61, stag=capi.HH, dom=dom)
    assert nt >= 4 and (o[:, 3:-3, 3:-3].sum(0) > 0).mean() > 0.9
    assert_bitexact(g, o, "LANDUSEF counts on E-grid mass points")
    dom = (orc.map_set(orc.PROJ_ROTLL, **kw), capi.map_set(capi.PROJ_ROTLL, **kw))
    g, o, nt = run_field(orc, handle, capi.SP_CONTINUOUS, "average_gcell(4.0)+four_pt+average_4pt", 1, 1, topo_tile, 120, 3, 0.00833333,
                         nx=39, ny=61, stag=capi.VV, dom=dom)
    assert np.abs(g - o).max() <= 1e-6 * np.abs(o).max(), "HGT_V on E-grid velocity points"


def test_sp_continuous_topography(orc, handle):
    """average_gcell(4.0)+four_pt+average_4pt on 30-arcsec integer topography with tile_bdr=3: sums of integers are
    exact in binary32, so the result is bit-exact despite the different summation order."""
    from wps_b200 import capi
    g, o, nt = run_field(orc, handle, capi.SP_CONTINUOUS, "average_gcell(4.0)+four_pt+average_4pt", 1, 1, topo_tile, 120, 3,
                         0.00833333, scale=None)
    assert_bitexact(g, o, "HGT_M")


def test_sp_continuous_multilevel_scaled_landmask(orc, handle):
    from wps_b200 import capi
    rng = np.random.default_rng(2)
    lm = (rng.random((56, 66)) < 0.8).astype(np.float32)
    g, o, _ = run_field(orc, handle, capi.SP_CONTINUOUS, "average_gcell(4.0)+four_pt+average_4pt", 3, 1, topo_tile, 100, 2,
                        0.00833333, landmask=lm, mask_val=0.0, fill=-999.0, scale=0.01)
    assert_bitexact(g, o, "3-level SP_CONTINUOUS")


@pytest.mark.parametrize("chain", ["four_pt", "sixteen_pt+four_pt+average_4pt+search(5)", "wt_average_4pt+search",
                                   "average_16pt+average_4pt"])
def test_continuous_pointwise(orc, handle, chain):
    """CONTINUOUS fields (e.g. 12 monthly albedo levels from a 0.144-deg dataset): per-point interp_sequence in each tile."""
    from wps_b200 import capi

    def fn(tx, ty, tn, nz):
        t = topo_tile(tx, ty, tn, nz)
        j, i = np.mgrid[ty:ty + tn, tx:tx + tn]
        t[:, (i * 3 + j) % 9 == 0] = -1e30
        return t
    g, o, _ = run_field(orc, handle, capi.CONTINUOUS, chain, 4, 1, fn, 50, 3, 0.144, nx=64, ny=48, dx=9000.0, msg=-1e30,
                        fill=-500.0)
    assert_bitexact(g, o, "continuous %s" % chain)


def test_raw_tile_decode(orc, handle):
    """wpsgpu_geogrid_add_tile_raw decodes read_geogrid.c byte layouts (big/little endian, signed, 1-4 bytes, row flip)
    identically to decoding on the host."""
    from wps_b200 import capi
    orc.set_transc_mode(1)
    odom, gdom = conus_like(30, 30, 9000.0, 30.0, 60.0, -98.0, 39.0, -105.0)
    xlat, xlon = orc.get_lat_lon_fields(odom, 1, 1, 30, 30, 1)
    kw = dict(dlat=0.01, dlon=0.01, known_x=1.0, known_y=1.0, known_lat=37.5, known_lon=-107.0)
    gsrc = capi.push_source_projection(capi.PROJ_LATLON, **kw)
    rng = np.random.default_rng(0)
    tn = 300
    for wordsize, signed, endian, row_order in ((1, 0, 0, 1), (2, 1, 0, 1), (2, 1, 1, 2), (3, 1, 0, 1), (4, 1, 1, 1), (1, 1, 0, 2)):
        hi = 1 << (8 * wordsize - (1 if signed else 0))
        lo = -hi if signed else 0
        vals = rng.integers(max(lo, -30000), min(hi, 30000), (1, tn, tn)).astype(np.int64)
        u = np.where(vals < 0, vals + (1 << (8 * wordsize)), vals).astype(np.uint64)
        stored = vals[:, ::-1, :] if row_order == 2 else vals
        us = np.where(stored < 0, stored + (1 << (8 * wordsize)), stored).astype(np.uint64)
        order = range(wordsize) if endian == 1 else range(wordsize - 1, -1, -1)
        raw = np.stack([((us >> (8 * b)) & 0xff).astype(np.uint8) for b in order], axis=-1).copy()
        dec = vals.astype(np.float32)
        if signed and wordsize < 4:  # read_geogrid.c: `ival > 2^(8w-1)` (strictly greater) => exactly 2^(8w-1) stays positive
            dec = np.where(u == (1 << (8 * wordsize - 1)), np.float32(1 << (8 * wordsize - 1)), dec).astype(np.float32)
        res = []
        for mode in ("raw", "float"):
            field = np.zeros((1, 30, 30), np.float32)
            handle.geogrid_field_begin(gdom, 1, xlat, xlon, 1, 1, 1, field, None)
            handle.geogrid_level_begin(gsrc, 1, capi.SP_CONTINUOUS, "average_gcell(1.0)+four_pt", 1e20, 1e20, -1.0)
            if mode == "raw":
                handle.geogrid_add_tile_raw(raw, (1, tn, tn), wordsize, signed, endian, row_order, 1, 1, 1, 0)
            else:
                handle.geogrid_add_tile(dec, 1, 1, 1, 0)
            handle.geogrid_level_end()
            handle.geogrid_field_finish()
            res.append(field)
        assert_bitexact(res[0], res[1], "raw decode wordsize=%d signed=%d endian=%d row_order=%d" % (wordsize, signed, endian, row_order))


@pytest.mark.parametrize("kind", ["categorical", "sp_continuous"])
def test_patch_sharded_field_equals_whole_domain(orc, handle, kind):
    """geogrid shards by destination patch (parallel_get_tile_dims) with the +-3 halo process_tile adds
    and no exchange step.  Four patches, each fed only the tiles that overlap it, reproduce the whole-domain field on
    their own window (halo included) -- which is what makes one-GPU-per-patch runs equal to a single-GPU run."""
    from wps_b200 import capi
    from wps_b200.shard import patch_bounds
    orc.set_transc_mode(1)
    nx, ny, dx, dxdeg = 60, 50, 9000.0, 0.00833333
    odom, gdom = conus_like(nx, ny, dx, 30.0, 60.0, -98.0, 39.0, -105.0)
    kw = dict(dlat=dxdeg, dlon=dxdeg, known_x=1.0, known_y=1.0, known_lat=-89.99583, known_lon=-179.99583)
    gsrc = capi.push_source_projection(capi.PROJ_LATLON, **kw)
    if kind == "categorical":
        itype, chain, nk, bdr, tile_fn, tile_n = capi.CATEGORICAL, "nearest_neighbor", 21, 0, cat_tile, 120
    else:
        itype, chain, nk, bdr, tile_fn, tile_n = capi.SP_CONTINUOUS, "average_gcell(4.0)+four_pt+average_4pt", 1, 3, topo_tile, 120

    def run(x0, x1, y0, y1):
        gx, gy = x1 - x0 + 1 + 6, y1 - y0 + 1 + 6
        xlat, xlon = orc.get_lat_lon_fields(odom, x0 - 3, y0 - 3, gx, gy, 1)
        dom_ll = (float(xlat.min()) - 0.3, float(xlat.max()) + 0.3, float(xlon.min()) - 0.3, float(xlon.max()) + 0.3)
        field = np.zeros((nk, gy, gx), np.float32)
        handle.geogrid_field_begin(gdom, 1, xlat, xlon, x0 - 3, y0 - 3, 1, field, None)
        handle.geogrid_level_begin(gsrc, 1, itype, chain, 1e20, 1e20, -1.0, None, 1, 1)
        tiles = tiles_for(dom_ll, dxdeg, tile_n, bdr)
        for (tx, ty) in tiles:
            handle.geogrid_add_tile(tile_fn(tx, ty, tile_n + 2 * bdr, 1 if itype == capi.CATEGORICAL else nk), tx, ty, 1, bdr)
        handle.geogrid_level_end()
        handle.geogrid_field_finish()
        return field, len(tiles)

    whole, nt_whole = run(1, nx, 1, ny)
    nts = []
    for rank in range(4):
        x0, x1, y0, y1 = patch_bounds(nx, ny, 4, rank)
        part, nt = run(x0, x1, y0, y1)
        nts.append(nt)
        ref = whole[:, y0 - 1:y1 + 6, x0 - 1:x1 + 6]     # whole-domain array starts at (-2, -2): patch window incl. halo
        # the outermost ring of any window only collects the source pixels on its inner half (the inside test of
        # where_maps_to, proc_point_module.F:296-303), so it is compared on the inner 2 halo cells + the patch itself
        part, ref = part[:, 1:-1, 1:-1], ref[:, 1:-1, 1:-1]
        if kind == "categorical":
            assert_bitexact(part, ref, "patch %d" % rank)
        else:
            np.testing.assert_allclose(part, ref, rtol=1e-6, atol=0.0)
    assert min(nts) <= nt_whole


@pytest.mark.parametrize("opt,npass", [(1, 2), (3, 1)])
def test_patchwise_smoothing_on_device(orc, handle, opt, npass):
    """a GPU that owns one destination patch smooths the patch extended by smoother_apron() cells and needs
    no halo exchange: on the patch the result equals the whole-domain smoothing bit for bit (oracle = whole domain)."""
    from wps_b200.shard import apron_window, patch_bounds, smoother_apron
    idim, jdim, world = 97, 71, 4
    a = np.random.default_rng(8).uniform(-50, 3000, (jdim, idim)).astype(np.float32)
    whole = orc.smooth(a, npass, opt)
    ap = smoother_apron(opt, npass)
    out = np.full_like(a, np.nan)
    for rank in range(world):
        x0, x1, y0, y1 = patch_bounds(idim, jdim, world, rank)
        ax0, ax1, ay0, ay1 = apron_window(x0, x1, y0, y1, idim, jdim, ap)
        w = np.ascontiguousarray(a[ay0 - 1:ay1, ax0 - 1:ax1])
        handle.smooth(w, opt, npass)
        out[y0 - 1:y1, x0 - 1:x1] = w[y0 - ay0:y1 - ay0 + 1, x0 - ax0:x1 - ax0 + 1]
    assert_bitexact(out, whole, "patch-wise smoothing opt=%d npass=%d" % (opt, npass))


def test_against_committed_geogrid_golden_vectors(handle):
    """tests/golden/geogrid_small.npz (oracle outputs, committed): land-use counts from four raw tiles, fractions /
    dominant category / landmask, the continuous accumulation with its point fallback, three smoothers and a Lambert wind
    rotation -- all through the C ABI, without the oracle at run time."""
    import os
    from wps_b200 import capi
    from tests.golden import make_golden_geogrid as mg
    g = np.load(os.path.join(os.path.dirname(mg.__file__), "geogrid_small.npz"))
    xlat, xlon = g["xlat"], g["xlon"]
    gy, gx = xlat.shape
    dom = capi.map_set(capi.PROJ_LC, **mg.KW_DOM)
    src = capi.push_source_projection(capi.PROJ_LATLON, **mg.KW_SRC)

    def field(itype, chain, nk, bdr, tile_fn, raw):
        f = np.zeros((nk, gy, gx), np.float32)
        handle.geogrid_field_begin(dom, capi.M, xlat, xlon, -2, -2, 1, f, None)
        handle.geogrid_level_begin(src, 1, itype, chain, 1e20, 1e20, -1.0, None, 1, 1)
        for tx, ty in mg.tile_origins(xlat, xlon, bdr):
            t = tile_fn(tx, ty, mg.TN + 2 * bdr)
            if raw:
                handle.geogrid_add_tile_raw(np.ascontiguousarray(t[0].astype(np.uint8)), t.shape, 1, 0, 0, 1, tx, ty, 1, bdr)
            else:
                handle.geogrid_add_tile(t, tx, ty, 1, bdr)
        handle.geogrid_level_end()
        handle.geogrid_field_finish(np.zeros((gy, capi.nwords(gx)), np.int32))
        return f

    counts = field(capi.CATEGORICAL, "nearest_neighbor", 21, 0, mg.cat_tile, True)
    assert_bitexact(counts, g["landuse_counts"], "golden land-use counts")
    fr = counts.copy()
    lm = np.zeros((gy, gx), np.float32)
    dc = np.zeros((gy, gx), np.float32)
    handle.fractions_dominant(fr, 1, 1e20, (17, 21), True, lm, dc)
    assert_bitexact(fr, g["landuse_fractions"], "golden fractions")
    assert_bitexact(dc, g["landuse_dominant"], "golden dominant category (landmask-aware, :699-727)")
    assert_bitexact(lm, g["landmask"], "golden landmask")
    fr2, dc2 = counts.copy(), np.zeros((gy, gx), np.float32)
    handle.fractions_dominant(fr2, 1, 1e20, (), True, None, dc2)
    assert_bitexact(dc2, g["landuse_dominant_plain"], "golden dominant category (plain, :1123-1144)")
    hgt = field(capi.SP_CONTINUOUS, "average_gcell(4.0)+four_pt+average_4pt", 1, 3, mg.topo_tile, False)
    # cell sums are binary64 on the device: the north-star tolerance for average_gcell, 1e-6 relative
    assert np.allclose(hgt, g["hgt"], rtol=1e-6, atol=0.0), "golden HGT"
    for name, opt, npass in (("hgt_121_2", 1, 2), ("hgt_smth_desmth_1", 2, 1), ("hgt_smth_desmth_special_2", 3, 2)):
        a = np.ascontiguousarray(g["hgt"][0] - np.float32(1600.0))
        handle.smooth(a, opt, npass)
        assert_bitexact(a, g[name], "golden smoother " + name)
    u, v = g["u"].copy()[None], g["v"].copy()[None]
    um = np.full((u.shape[1], capi.nwords(u.shape[2])), -1, np.int32)
    vm = np.full((v.shape[1], capi.nwords(v.shape[2])), -1, np.int32)
    handle.rotate_winds(1, dom, u, um, v, vm, g["xlon_u"], g["xlon_v"])
    assert ulp_diff(u[0], g["u_rot"]).max() <= 1 and ulp_diff(v[0], g["v_rot"]).max() <= 1, "golden wind rotation"


@pytest.mark.parametrize("dxdeg,tn", [(0.01, 300), (0.3, 24)])
def test_raw_categorical_tiles(orc, handle, dxdeg, tn):
    """Categorical tiles handed over as file bytes: the batched accumulate kernel reads the bytes itself and the tile is only
    decoded for the point pass (k_geo_decode_b behind a gate).  Every byte layout of read_geogrid.c, for a source finer than
    the domain (no point-pass candidates: the gate stays shut) and a much coarser one (most points take the chain from the
    lazily decoded tile): identical to the same tiles passed as decoded binary32 values, counts and invalid-category counter."""
    from wps_b200 import capi
    orc.set_transc_mode(1)
    odom, gdom = conus_like(30, 30, 9000.0, 30.0, 60.0, -98.0, 39.0, -105.0)
    xlat, xlon = orc.get_lat_lon_fields(odom, 1, 1, 30, 30, 1)
    kw = dict(dlat=dxdeg, dlon=dxdeg, known_x=1.0, known_y=1.0, known_lat=37.5 - (3.0 if dxdeg > 0.1 else 0.0), known_lon=-107.0 - (3.0 if dxdeg > 0.1 else 0.0))
    gsrc = capi.push_source_projection(capi.PROJ_LATLON, **kw)
    rng = np.random.default_rng(5)
    for wordsize, signed, endian, row_order in ((1, 0, 0, 1), (1, 1, 0, 2), (2, 1, 0, 1), (2, 0, 1, 2), (3, 1, 1, 1), (4, 1, 0, 2)):
        vals = rng.integers(1, 22, (1, tn, tn)).astype(np.int64)
        vals[0, ::7, ::5] = 30                                   # invalid categories (WARN counter, proc_point_module.F:348)
        if signed:
            vals[0, 1::9, 2::4] = -3
        stored = vals[:, ::-1, :] if row_order == 2 else vals
        us = np.where(stored < 0, stored + (1 << (8 * wordsize)), stored).astype(np.uint64)
        order = range(wordsize) if endian == 1 else range(wordsize - 1, -1, -1)
        raw = np.stack([((us >> (8 * b)) & 0xff).astype(np.uint8) for b in order], axis=-1).copy()
        dec = vals.astype(np.float32)
        res, ninv = [], []
        for mode in ("raw", "float"):
            field = np.zeros((21, 30, 30), np.float32)
            handle.geogrid_field_begin(gdom, 1, xlat, xlon, 1, 1, 1, field, None)
            handle.geogrid_level_begin(gsrc, 1, capi.CATEGORICAL, "nearest_neighbor", 1e20, 1e20, -1.0)
            if mode == "raw":
                handle.geogrid_add_tile_raw(raw, (1, tn, tn), wordsize, signed, endian, row_order, 1, 1, 1, 0)
            else:
                handle.geogrid_add_tile(dec, 1, 1, 1, 0)
            handle.geogrid_level_end()
            ninv.append(handle.geogrid_field_finish())
            res.append(field)
        what = "raw categorical wordsize=%d signed=%d endian=%d row_order=%d dx=%g" % (wordsize, signed, endian, row_order, dxdeg)
        assert_bitexact(res[0], res[1], what)
        assert ninv[0] == ninv[1], what
        assert res[0].sum() > 0, what
