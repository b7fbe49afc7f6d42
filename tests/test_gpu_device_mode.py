"""GPU parity in DEVICE-pointer mode (wpsgpu_set_pointer_mode / Handle(dev, PTR_DEVICE)): the mode bench.py times.
Inputs and outputs are torch tensors resident in HBM; results must equal the oracle (and hence the host-pointer mode)."""
import numpy as np
import pytest

from tests.util import assert_bitexact, assert_sixteen_pt_tolerance, bool_to_bits, conus_like, halo, landsea, latlon_source, synth_field

pytestmark = pytest.mark.gpu


@pytest.fixture()
def dhandle():
    from wps_b200 import capi
    h = capi.Handle(0, capi.PTR_DEVICE)
    yield h
    h.close()


@pytest.mark.parametrize("tolerance", [False, True])
def test_metgrid_137_levels_device_resident(orc, dhandle, tolerance):
    """(tolerance = True: the same through the tolerance-mode strip kernel, job groups of 64 + 64 + 9 levels.)
    Config-5 shape in small: 137 levels of one field (three job groups of at most 64 slabs in device mode) plus a
    masked soil chain with deferred searches, all device-resident, one flush; oracle on a spread of levels."""
    import torch
    from wps_b200 import capi
    dhandle.metgrid_set_precision(capi.PRECISION_TOLERANCE if tolerance else capi.PRECISION_EXACT)
    dev = torch.device("cuda", 0)
    nx, ny, nlev = 150, 110, 137
    orc.set_transc_mode(1)
    odom, gdom = conus_like(nx, ny, 12000.0, truelat1=30.0, truelat2=60.0, stand_lon=-98.0, ref_lat=34.83, ref_lon=-81.03)
    xlat, xlon = orc.get_lat_lon_fields(odom, 1, 1, nx, ny, orc.M)
    osrc, gsrc = latlon_source(720, 361, -0.5, 0.5, 90.0, 0.0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_xlat, d_xlon = t(xlat), t(xlon)
    rng = np.random.default_rng(9)
    landmask = (rng.random((ny, nx)) < 0.6).astype(np.float32)
    d_lm = t(landmask)
    dhandle.metgrid_set_grid(0, d_xlat, d_xlon, 1, 1)
    dhandle.metgrid_set_landmask(0, d_lm)
    dhandle.metgrid_set_domain(1, nx, 1, ny)
    slabs = np.stack([halo(synth_field(720, 361, 2000 + s, kind="wind" if s % 3 else "smooth", zeros=0.002)) for s in range(nlev)])
    d_slabs = t(slabs)
    d_r = torch.zeros((nlev, ny, nx), dtype=torch.float32, device=dev)
    d_np = torch.zeros((nlev, ny, capi.nwords(nx)), dtype=torch.int32, device=dev)
    chain = "sixteen_pt+four_pt+average_4pt"
    fo_g, fo_o = capi.field_opts(chain, 1e20, 0.0, -2.0), orc.field_opts(chain, 1e20, 0.0, -2.0)
    descs = (capi.Slab * nlev)()
    for l in range(nlev):
        capi.Handle.slab_desc(d_slabs[l], -2, 1, 3, 0, capi.M, d_r[l], d_np[l], out=descs[l])
    dhandle.metgrid_enqueue_slabs(gsrc, fo_g, descs)
    # masked soil field: missing over water, interp_mask = LANDSEA(0), masked = water, search at the end of the chain
    ls = landsea(720, 361)
    soil = synth_field(720, 361, 3000)
    soil[ls == 0] = np.float32(-1e30)
    d_soil, d_mask = t(halo(soil)), t(halo(ls))
    d_rs = torch.zeros((ny, nx), dtype=torch.float32, device=dev)
    d_ns = torch.zeros((ny, capi.nwords(nx)), dtype=torch.int32, device=dev)
    schain = "sixteen_pt+four_pt+wt_average_4pt+wt_average_16pt+search"
    fs_g = capi.field_opts(schain, np.float32(-1e30), 285.0, 0.0, (0.0, 1e20, 1e20))
    fs_o = orc.field_opts(schain, np.float32(-1e30), 285.0, 0.0, (0.0, 1e20, 1e20))
    dhandle.metgrid_enqueue_slab(gsrc, fs_g, d_soil, -2, 1, 3, 0, capi.M, d_rs, d_ns, masks=(d_mask, None, None), use_landmask=True)
    dhandle.metgrid_flush()
    dhandle.synchronize()
    r, npw = d_r.cpu().numpy(), d_np.cpu().numpy()
    for l in (0, 1, 63, 64, 65, 127, 128, 136):
        r_ref, np_ref = orc.interp_met_field(osrc, fo_o, 1, xlat, xlon, 1, 1, (1, nx, 1, ny), slabs[l], -2, 1, 3)
        assert (npw[l] == np_ref).all(), "level %d" % l
        if tolerance:
            assert_sixteen_pt_tolerance(r[l], r_ref, float(np.abs(slabs[l]).max()), "level %d, tolerance mode" % l)
        else:
            assert_bitexact(r[l], r_ref, "level %d" % l)
    r_ref, np_ref = orc.interp_met_field(osrc, fs_o, 1, xlat, xlon, 1, 1, (1, nx, 1, ny), halo(soil), -2, 1, 3, mask=halo(ls),
                                         landmask=landmask)
    assert (d_ns.cpu().numpy() == np_ref).all()
    if tolerance:
        assert_sixteen_pt_tolerance(d_rs.cpu().numpy(), r_ref, float(np.abs(soil[ls != 0]).max()), "masked soil chain, tolerance mode")
    else:
        assert_bitexact(d_rs.cpu().numpy(), r_ref, "masked soil chain, device mode")


def test_vertical_fill_and_raw_records_device_resident(orc, dhandle):
    """fill_missing_levels and the raw-record ingest with every array in HBM (pointer tables stay host arrays)."""
    import torch
    from wps_b200 import capi
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(12)
    nx, ny, nlev = 90, 37, 10
    levels = [1000 * (k + 1) for k in range(nlev)]
    r = rng.uniform(0, 1, (nlev, ny, nx)).astype(np.float32)
    v = np.stack([bool_to_bits(rng.random((ny, nx)) < 0.55) for _ in range(nlev)])
    r_ref, v_ref = orc.fill_missing_levels(levels, r, v)
    d_r, d_v = torch.from_numpy(r).to(dev), torch.from_numpy(v).to(dev)
    dhandle.fill_missing_levels(levels, [d_r[k] for k in range(nlev)], [d_v[k] for k in range(nlev)])
    dhandle.synchronize()
    assert (d_v.cpu().numpy() == v_ref).all()
    assert_bitexact(d_r.cpu().numpy(), r_ref, "fill_missing_levels, device mode")
    # raw big-endian record resident on the device
    gx, gy = 61, 47
    orc.set_transc_mode(1)
    odom, _ = conus_like(gx, gy, 30000.0, truelat1=30.0, truelat2=60.0, stand_lon=-98.0, ref_lat=34.83, ref_lon=-81.03)
    xlat, xlon = orc.get_lat_lon_fields(odom, 1, 1, gx, gy, orc.M)
    osrc, gsrc = latlon_source(360, 181, -1.0, 1.0, 90.0, 0.0)
    rec = synth_field(360, 181, 5)
    d_raw = torch.from_numpy(rec.astype(">f4").view(np.float32).copy()).to(dev)
    d_out = torch.zeros((gy, gx), dtype=torch.float32, device=dev)
    d_bits = torch.zeros((gy, capi.nwords(gx)), dtype=torch.int32, device=dev)
    dhandle.metgrid_set_grid(0, torch.from_numpy(xlat).to(dev), torch.from_numpy(xlon).to(dev), 1, 1)
    dhandle.metgrid_set_domain(1, gx, 1, gy)
    fo_g, fo_o = capi.field_opts("sixteen_pt+four_pt+average_4pt", 1e20, 0.0, -2.0), orc.field_opts("sixteen_pt+four_pt+average_4pt", 1e20, 0.0, -2.0)
    dhandle.metgrid_enqueue_slab(gsrc, fo_g, d_raw, -2, 1, 3, 0, capi.M, d_out, d_bits, raw_big_endian=True)
    dhandle.metgrid_flush()
    dhandle.synchronize()
    r_ref, np_ref = orc.interp_met_field(osrc, fo_o, 1, xlat, xlon, 1, 1, (1, gx, 1, gy), halo(rec), -2, 1, 3)
    assert (d_bits.cpu().numpy() == np_ref).all()
    assert_bitexact(d_out.cpu().numpy(), r_ref, "raw record, device mode")


@pytest.mark.parametrize("kind", ["categorical", "continuous"])
def test_geogrid_tiles_pipelined_in_device_mode(handle, dhandle, kind):
    """Device-pointer mode stages tile k+1 (decode + where_maps_to) on a second stream while tile k is accumulated, merged
    and point-filled (two buffer sets, geogrid.cu geo_tile_begin).  The field must equal, bit for bit, what the
    host-pointer mode (one tile at a time on one stream; checked against the oracle in test_gpu_geogrid.py) produces from
    the same tiles -- for a categorical field fed as raw bytes and a continuous one fed as float tiles with a border."""
    import torch
    from wps_b200 import capi
    from tests.test_gpu_geogrid import cat_tile, tiles_for, topo_tile
    dev = torch.device("cuda", 0)
    nx, ny, dx, dxdeg = 90, 70, 9000.0, 0.00833333
    _, gdom = conus_like(nx, ny, dx, 30.0, 60.0, -98.0, 39.0, -105.0)
    gx, gy = nx + 6, ny + 6
    xlat, xlon = handle.xytoll_grid(gdom, capi.M, -2, -2, gx, gy)
    kw = dict(dlat=dxdeg, dlon=dxdeg, known_x=1.0, known_y=1.0, known_lat=-89.99583, known_lon=-179.99583)
    gsrc = capi.push_source_projection(capi.PROJ_LATLON, **kw)
    dom_ll = (float(xlat.min()) - 0.3, float(xlat.max()) + 0.3, float(xlon.min()) - 0.3, float(xlon.max()) + 0.3)
    if kind == "categorical":
        itype, chain, nk, bdr, tn, msg = capi.CATEGORICAL, "nearest_neighbor", 21, 0, 100, 255.0
    else:
        itype, chain, nk, bdr, tn, msg = capi.SP_CONTINUOUS, "average_gcell(4.0)+four_pt+average_4pt", 1, 3, 100, 1e20
    tiles = tiles_for(dom_ll, dxdeg, tn, bdr)
    assert len(tiles) >= 6
    tsz = tn + 2 * bdr
    data = [(cat_tile(tx, ty, tsz, 1) if kind == "categorical" else topo_tile(tx, ty, tsz, 1)) for tx, ty in tiles]
    init = np.random.default_rng(4).uniform(-1, 1, (nk, gy, gx)).astype(np.float32)

    # host-pointer mode, float tiles
    f_host = init.copy()
    handle.geogrid_field_begin(gdom, capi.M, xlat, xlon, -2, -2, 1, f_host, None)
    handle.geogrid_level_begin(gsrc, 1, itype, chain, msg, 1e20, -1.0, None, 1, 1)
    for (tx, ty), t in zip(tiles, data):
        handle.geogrid_add_tile(t, tx, ty, 1, bdr)
    handle.geogrid_level_end()
    p_host = np.zeros((gy, capi.nwords(gx)), np.int32)
    handle.geogrid_field_finish(p_host)

    # device-pointer mode: raw bytes (categorical) / float tiles (continuous), everything resident in HBM
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_xlat, d_xlon, d_field = t(xlat), t(xlon), t(init)
    d_tiles = [t(x[0].astype(np.uint8)) if kind == "categorical" else t(x) for x in data]
    torch.cuda.synchronize()
    dhandle.geogrid_field_begin(gdom, capi.M, d_xlat, d_xlon, -2, -2, 1, d_field)
    dhandle.geogrid_level_begin(gsrc, 1, itype, chain, msg, 1e20, -1.0, None, 1, 1)
    for (tx, ty), dt in zip(tiles, d_tiles):
        if kind == "categorical":
            dhandle.geogrid_add_tile_raw(dt, (1, tsz, tsz), 1, 0, 0, 1, tx, ty, 1, bdr)
        else:
            dhandle.geogrid_add_tile(dt, tx, ty, 1, bdr)
    dhandle.geogrid_level_end()
    dhandle.geogrid_field_finish()
    dhandle.synchronize()
    assert_bitexact(d_field.cpu().numpy(), f_host, "geogrid field, device mode (pipelined tiles) vs host mode, %s" % kind)


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 2), (1, 33, 70), (3, 67, 131)])
@pytest.mark.parametrize("opt,npass", [(1, 1), (1, 4), (2, 1), (2, 2), (3, 1), (3, 3)])
def test_smoothers_device_resident_all_shapes(orc, dhandle, shape, opt, npass):
    """The fused smoother pass (one launch per pass, tile + apron in shared memory) with the array resident in HBM: a
    single pass goes through a scratch array and is copied back, several passes ping-pong and the last one writes in
    place.  Degenerate shapes (no interior points) must come back unchanged.  Bit for bit against the oracle."""
    import torch
    rng = np.random.default_rng(opt * 100 + npass * 10 + shape[2])
    a = rng.uniform(-300, 3000, shape).astype(np.float32)
    d = torch.from_numpy(a).to(torch.device("cuda", 0))
    dhandle.smooth(d, opt, npass)
    dhandle.synchronize()
    assert_bitexact(d.cpu().numpy(), orc.smooth(a, npass, opt).reshape(shape), "smooth opt=%d npass=%d shape=%s" % (opt, npass, shape))
