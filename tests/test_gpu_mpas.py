"""GPU parity of the MPAS remapping path (wps_b200/csrc/mpas.cu, remapper.F) through the C ABI, against oracle/orc_mpas.c."""
import numpy as np
import pytest

from tests.util import ulp_diff
from wps_b200 import synth

pytestmark = pytest.mark.gpu

KEYS = {0: ("nearestCell", "sourceCells", "cellWeights"), 1: ("nearestVertex", "sourceVertices", "vertexWeights"),
        2: ("nearestEdge", "sourceEdges", "edgeWeights"), 3: (None, "sourceMaskedCells", "cellMaskedWeights")}


def assert_weights(w, ref, what):
    """Weights are bit-identical wherever the binary64 sin / cos / tan / asin / atan of CUDA and glibc round to the same
    binary32 value (the shared FP64-then-round policy, as for the projections).  A 1-ulp difference in one of them is amplified
    by the cancellation in Heron's formula on small triangles -- binary32 in the reference too, whose own weights move by up to
    4e-4 (cells) between glibc's float and double functions -- and without bound where the 'polygon' of a cell's first three
    vertices / edges (:204, :238) is nearly degenerate and the weights reach 1e2..1e3: those points are compared for being
    finite at the same places only."""
    assert np.array_equal(np.isfinite(w), np.isfinite(ref)), what
    same = (w.view(np.int32) == ref.view(np.int32)) | ~np.isfinite(ref)
    assert same.mean() > 0.999, (what, same.mean())
    ok = np.isfinite(ref).all(-1) & (np.abs(np.nan_to_num(ref)).max(-1) < 50.0)
    d = np.abs(w[ok].astype(np.float64) - ref[ok]) / np.maximum(1.0, np.abs(ref[ok]))
    assert d.max() < 1e-3, (what, d.max())


def lambert_target(nx, ny, dx, stagger=None):
    """lat / lon in radians of a Lambert grid over North America (process_mpas_fields: xlat * deg2rad, :1538-1541)."""
    from wps_b200 import capi
    dom = synth.lambert_domain(nx, ny, dx, 30.0, 60.0, -98.0, 34.83, -81.03)
    h = capi.Handle(0)
    xlat, xlon = h.xytoll_grid(dom, stagger or capi.M, 1, 1, nx, ny)
    h.close()
    d2r = np.float32(np.arcsin(np.float32(1.0)) / np.float32(90.0))
    return np.ascontiguousarray(xlat * d2r), np.ascontiguousarray(xlon * d2r)


@pytest.fixture(scope="module")
def setup(orc):
    from wps_b200 import capi
    mesh = synth.mpas_mesh(10242, seed=2)
    lat, lon = lambert_target(150, 120, 60000.0)
    orc.set_transc_mode(1)
    t = orc.mpas_remap_setup(mesh, lat, lon)
    h = capi.Handle(0)
    h.mpas_set_mesh(mesh)
    h.mpas_remap_setup(0, lat, lon)
    yield mesh, lat, lon, t, h
    h.close()


def test_tables_equal_the_sequential_scan(setup):
    """Every table of remap_info_setup, bit for bit: the device reproduces the reference's point-after-point scan (each walk
    starts at the previous point's result), including the masked scan whose start is a cell index after a search point."""
    mesh, lat, lon, t, h = setup
    nsearch = 0
    for kind, (kn, ks, kw) in KEYS.items():
        near, src, w = h.mpas_get_tables(0, kind)
        if kn:
            assert np.array_equal(near, t[kn]), kn
        assert np.array_equal(src, t[ks]), ks
        assert_weights(w, t[kw], kw)
        if kind == 3:
            nsearch = int(((src[..., 1] == 1) & (src[..., 2] == 1)).sum())
    assert nsearch > 1000            # open-ocean points went through search_for_cells
    # start sensitivity is real on this mesh: the edge walk has points with a second fixed point
    nodes = mesh["edgesOnCell"][t["nearestCell"] - 1]
    assert ((nodes == t["nearestEdge"][..., None]).any(-1)).mean() < 1.0


def test_second_grid_and_other_meshes(orc):
    """U-staggered target on grid slot 1, a coarse and a finer mesh, a target that crosses the date line and the pole."""
    from wps_b200 import capi
    orc.set_transc_mode(1)
    h = capi.Handle(0)
    for ncells, seed in ((642, 5), (40962, 6)):
        mesh = synth.mpas_mesh(ncells, seed=seed)
        h.mpas_set_mesh(mesh)
        lat, lon = lambert_target(91, 70, 90000.0, stagger=capi.U)
        ny, nx = 40, 64
        lat2 = np.ascontiguousarray(np.deg2rad(np.linspace(60.0, 89.9, ny, dtype=np.float32))[:, None] * np.ones((1, nx), np.float32))
        lon2 = np.ascontiguousarray(np.deg2rad(np.linspace(-180.0, 180.0, nx, dtype=np.float32))[None, :] * np.ones((ny, 1), np.float32))
        for slot, (la, lo) in ((1, (lat, lon)), (2, (lat2, lon2))):
            t = orc.mpas_remap_setup(mesh, la, lo)
            h.mpas_remap_setup(slot, la, lo)
            for kind, (kn, ks, kw) in KEYS.items():
                near, src, w = h.mpas_get_tables(slot, kind)
                if kn:
                    assert np.array_equal(near, t[kn]), (ncells, slot, kn)
                assert np.array_equal(src, t[ks]), (ncells, slot, ks)
                assert_weights(w, t[kw], (ncells, slot, kw))
    h.close()


def test_all_water_and_all_land_masks(orc):
    """landmask all 0: every point runs search_for_cells to exhaustion (cells (1,1,1), weights 0) -- with the bounded queue of
    2700 entries overrunning on the 10242-cell mesh; all 1: no search at all."""
    from wps_b200 import capi
    orc.set_transc_mode(1)
    h = capi.Handle(0)
    mesh = synth.mpas_mesh(10242, seed=2)
    lat, lon = lambert_target(40, 30, 150000.0)
    for val in (0, 1):
        mesh["landmask"][:] = val
        h.mpas_set_mesh(mesh)
        h.mpas_remap_setup(0, lat, lon)
        t = orc.mpas_remap_setup(mesh, lat, lon)
        _, src, w = h.mpas_get_tables(0, capi.MPAS_MASKED_CELLS)
        assert np.array_equal(src, t["sourceMaskedCells"])
        assert_weights(w, t["cellMaskedWeights"], val)
        if val == 0:
            assert (src == 1).all() and (w == 0.0).all()
    h.close()


def device_tables(h, t):
    """the oracle's table dict with the device's weights (identical integer tables are asserted elsewhere)"""
    from wps_b200 import capi
    d = dict(t)
    for kind, (kn, ks, kw) in KEYS.items():
        _, src, w = h.mpas_get_tables(0, kind)
        assert np.array_equal(src, t[ks])
        d[kw] = w
    return d


def test_remap_field_bit_exact(setup, orc):
    from wps_b200 import capi
    mesh, lat, lon, t, h = setup
    t = device_tables(h, t)
    for nlev in (1, 4, 55):
        f = synth.mpas_field(mesh, nlev, seed=nlev)
        for masked, ks, kw in ((False, "sourceCells", "cellWeights"), (True, "sourceMaskedCells", "cellMaskedWeights")):
            got = h.mpas_remap_field(0, capi.MPAS_CELLS, f, masked=masked, nw=3)
            ref = orc.mpas_remap_field(t[ks], t[kw], f, nw=3)
            assert np.array_equal(got.view(np.int32), ref.view(np.int32)), (nlev, masked)
    # 1-d sources on vertices / edges: all maxEdges table entries take part (remapper.F:506-511)
    for decomp, where, ks, kw in ((capi.MPAS_VERTICES, "Vertex", "sourceVertices", "vertexWeights"), (capi.MPAS_EDGES, "Edge", "sourceEdges", "edgeWeights")):
        v = synth.mpas_field(mesh, 1, seed=9, where=where)[:, 0].copy()
        got = h.mpas_remap_field(0, decomp, v)
        ref = orc.mpas_remap_field(t[ks], t[kw], v)
        assert np.array_equal(got.view(np.int32), ref.view(np.int32)), where
        v3 = synth.mpas_field(mesh, 3, seed=10, where=where)
        got = h.mpas_remap_field(0, decomp, v3)
        ref = orc.mpas_remap_field(t[ks], t[kw], v3)
        assert np.array_equal(got.view(np.int32), ref.view(np.int32)), where
    # DOUBLE (REAL weight promoted) and INTEGER (nearest node) sources
    f64 = synth.mpas_field(mesh, 5, seed=11).astype(np.float64) * 1.0000001234
    got = h.mpas_remap_field(0, capi.MPAS_CELLS, f64)
    assert np.array_equal(got.view(np.int64), orc.mpas_remap_field(t["sourceCells"], t["cellWeights"], f64).view(np.int64))
    for decomp, where, kn in ((capi.MPAS_CELLS, "Cell", "nearestCell"), (capi.MPAS_VERTICES, "Vertex", "nearestVertex"), (capi.MPAS_EDGES, "Edge", "nearestEdge")):
        fi = (synth.mpas_field(mesh, 2, seed=12, where=where) * 100).astype(np.int32)
        assert np.array_equal(h.mpas_remap_field(0, decomp, fi), orc.mpas_remap_field_int(t[kn], fi)), where


def test_remap_planes_fused_extraction(setup, orc):
    """remap + per-level planes of process_mpas_fields: plain levels, the nVertLevelsP1 rule, landmask fill of masked=water
    fields (process_domain_module.F:1700-1956), more levels than one launch takes (256)."""
    from wps_b200 import capi
    mesh, lat, lon, t, h = setup
    t = device_tables(h, t)
    ny, nx = lat.shape
    lm = (np.sin(np.arange(ny * nx, dtype=np.float64) * 0.37).reshape(ny, nx) > -0.2).astype(np.float32)
    for nlev, mode, masked, landmask in ((55, 0, False, None), (56, 1, False, None), (4, 0, True, lm), (1, 0, True, lm), (300, 1, False, None)):
        f = synth.mpas_field(mesh, nlev, seed=20 + nlev)
        ks, kw = ("sourceMaskedCells", "cellMaskedWeights") if masked else ("sourceCells", "cellWeights")
        wrf = orc.mpas_remap_field(t[ks], t[kw], f, nw=3)
        ref = orc.mpas_extract_planes(wrf, mode, landmask=landmask, fill=-1.0e30)
        planes = [np.full((ny, nx), np.nan, np.float32) for _ in range(nlev)]
        h.mpas_remap_planes(0, capi.MPAS_CELLS, f if nlev > 1 else f[:, 0].copy(), planes, mode=mode, masked=masked, landmask=landmask, fill_missing=-1.0e30)
        got = np.stack(planes)
        assert np.array_equal(got.view(np.int32), ref.view(np.int32)), (nlev, mode, masked)


def test_device_pointer_mode(setup, orc):
    import torch
    from wps_b200 import capi
    mesh, lat, lon, t, h0 = setup
    t = device_tables(h0, t)
    h = capi.Handle(0, capi.PTR_DEVICE)
    dev = torch.device("cuda", 0)
    h.mpas_set_mesh(mesh)                                         # the mesh is always a host structure
    h.mpas_remap_setup(0, torch.from_numpy(lat).to(dev), torch.from_numpy(lon).to(dev))
    f = synth.mpas_field(mesh, 8, seed=30)
    fd = torch.from_numpy(f).to(dev)
    out = h.mpas_remap_field(0, capi.MPAS_CELLS, fd)
    torch.cuda.synchronize()
    ref = orc.mpas_remap_field(t["sourceCells"], t["cellWeights"], f, nw=3)
    assert np.array_equal(out.cpu().numpy().view(np.int32), ref.view(np.int32))
    planes = [torch.empty(lat.shape, dtype=torch.float32, device=dev) for _ in range(8)]
    h.mpas_remap_planes(0, capi.MPAS_CELLS, fd, planes)
    h.synchronize()
    assert np.array_equal(torch.stack(planes).cpu().numpy().view(np.int32), np.moveaxis(ref, 2, 0).view(np.int32))
    h.close()


def test_errors(setup):
    from wps_b200 import capi
    mesh, lat, lon, t, _ = setup
    h = capi.Handle(0)
    with pytest.raises(capi.WpsGpuError, match="no MPAS mesh"):
        h._mpas_grids = {}
        h.mpas_remap_setup(0, lat, lon)
    bad = dict(mesh)
    bad["cellsOnCell"] = mesh["cellsOnCell"].copy()
    bad["cellsOnCell"][5, 0] = mesh["nCells"] + 7
    with pytest.raises(capi.WpsGpuError, match="limited-area|out of range"):
        h.mpas_set_mesh(bad)
    h.mpas_set_mesh(mesh)
    with pytest.raises(capi.WpsGpuError, match="has not run"):
        h._mpas_grids[1] = lat.shape
        h.mpas_remap_field(1, capi.MPAS_CELLS, synth.mpas_field(mesh, 2))
    with pytest.raises(capi.WpsGpuError, match="grid_id"):
        h._mpas_grids[5] = lat.shape
        h.mpas_remap_field(5, capi.MPAS_CELLS, synth.mpas_field(mesh, 2))
    h.close()


TABLE = """
========================================
name=TT ; mpas_name=theta ; output_stagger=M
========================================
name=PRESSURE ; mpas_name=pressure
========================================
name=UU ; mpas_name=uReconstructZonal ; output_stagger=U ; is_u_field=yes
========================================
name=GHT ; mpas_name=zgrid
========================================
name=SKINTEMP ; mpas_name=skintemp ; masked=both ; fill_missing=0.
========================================
name=ST ; mpas_name=tslb ; masked=water ; fill_missing=285.
========================================
name=SM ; mpas_name=smois ; masked=water ; fill_missing=1.
========================================
name=SEAICE ; mpas_name=xice ; masked=water ; fill_missing=0.
========================================
"""


def test_process_mpas_fields_host_mirror(orc):
    """The field loop of process_mpas_fields (process_domain_module.F:1626-1958) above the ABI: table lookup by mpas_name,
    renames, target grid by output stagger, masked table for masked=water fields on the mass grid, the (field, level) planes
    of every dimension case -- values against the oracle's remap + extraction."""
    from wps_b200 import capi, interp_option, mpas
    orc.set_transc_mode(1)
    mesh = synth.mpas_mesh(2562, seed=8)
    entries = interp_option.read_interp_table(TABLE)
    h = capi.Handle(0)
    h.mpas_set_mesh(mesh)
    grids = {}
    for stag, gid in ((capi.M, 0), (capi.U, 1)):
        lat, lon = lambert_target(60 + (stag == capi.U), 50, 120000.0, stagger=stag)
        h.mpas_remap_setup(gid, lat, lon)
        t = orc.mpas_remap_setup(mesh, lat, lon)
        for kind, (kn, ks, kw) in KEYS.items():
            t[kw] = h.mpas_get_tables(gid, kind)[2]
        grids[gid] = t
    ny, nx = 50, 60
    lm = (np.cos(np.arange(ny * nx) * 0.11).reshape(ny, nx) > 0).astype(np.float32)

    def ref_planes(gid, f, mode=0, masked=False, landmask=None, fill=0.0):
        t = grids[gid]
        ks, kw = ("sourceMaskedCells", "cellMaskedWeights") if masked else ("sourceCells", "cellWeights")
        return orc.mpas_extract_planes(orc.mpas_remap_field(t[ks], t[kw], f, nw=3), mode, landmask=landmask, fill=fill)

    def same(a, b):
        return np.array_equal(a.view(np.int32), b.view(np.int32))
    th = synth.mpas_field(mesh, 6, seed=1)
    r = mpas.process_mpas_field(h, entries, "theta", ("nVertLevels", "nCells"), th, landmask=lm)
    assert [(d["field"], d["level"], d["vertical_coord"]) for d in r] == [("TT", k, "num_mpas_levels") for k in range(1, 7)]
    assert same(np.stack([d["r_arr"] for d in r]), ref_planes(0, th))
    # derive_mpas_fields (:1957-2053): with PRESSURE on the same levels, TT = theta * (p / 1e5) ** (287 / 1004), in place
    pr = (np.float32(1.0e5) * np.exp(-np.arange(6, dtype=np.float32) / 5.0)[None, :] * (1.0 + 0.01 * synth.mpas_field(mesh, 6, seed=7) / 300.0)).astype(np.float32)
    rp = mpas.process_mpas_field(h, entries, "pressure", ("nVertLevels", "nCells"), pr)
    theta_planes = np.stack([d["r_arr"] for d in r]).copy()
    mpas.derive_mpas_fields(h, r + rp)
    for k in range(6):
        want = orc.mpas_derive_tt(theta_planes[k], rp[k]["r_arr"])
        assert ulp_diff(r[k]["r_arr"], want).max() <= 1, k
        ex = (rp[k]["r_arr"].astype(np.float64) / 1.0e5) ** (287.0 / 1004.0)
        assert np.abs(r[k]["r_arr"] / (theta_planes[k].astype(np.float64) * ex) - 1.0).max() < 1e-6, k
    mpas.derive_mpas_fields(h, r)                      # no PRESSURE levels: nothing happens (:1998)
    # t2m is renamed to theta (:1634): a 1-d field stored at 200100
    t2 = np.ascontiguousarray(th[:, 0])
    r = mpas.process_mpas_field(h, entries, "t2m", ("nCells",), t2)
    assert [(d["field"], d["level"], d["vertical_coord"]) for d in r] == [("TT", 200100.0, "none")]
    assert same(r[0]["r_arr"][None], ref_planes(0, t2))
    # u10 -> uReconstructZonal -> the U grid
    r = mpas.process_mpas_field(h, entries, "u10", ("nCells",), t2)
    assert r[0]["stagger"] == capi.U and r[0]["r_arr"].shape == (50, 61) and same(r[0]["r_arr"][None], ref_planes(1, t2))
    # zgrid on nVertLevelsP1: surface plane, interface means, and SOILHGT (:1733-1809)
    zg = synth.mpas_field(mesh, 5, seed=2)
    r = mpas.process_mpas_field(h, entries, "zgrid", ("nVertLevelsP1", "nCells"), zg)
    assert [(d["field"], d["level"]) for d in r] == [("GHT", 200100.0), ("GHT", 1), ("GHT", 2), ("GHT", 3), ("GHT", 4), ("SOILHGT", 200100.0)]
    assert same(np.stack([d["r_arr"] for d in r[:5]]), ref_planes(0, zg, mode=1)) and r[5]["r_arr"] is r[0]["r_arr"]
    # soil fields: names per layer, masked table, landmask fill with the entry's fill_missing (:1811-1925)
    ts = synth.mpas_field(mesh, 4, seed=3)
    r = mpas.process_mpas_field(h, entries, "tslb", ("nSoilLevels", "nCells"), ts, landmask=lm)
    assert [d["field"] for d in r] == ["ST000010", "ST010040", "ST040100", "ST100200"] and all(d["level"] == 200100.0 for d in r)
    assert same(np.stack([d["r_arr"] for d in r]), ref_planes(0, ts, masked=True, landmask=lm, fill=285.0))
    assert (np.stack([d["r_arr"] for d in r])[:, lm == 0] == 285.0).all()
    with pytest.raises(capi.WpsGpuError, match="Too many soil layers"):
        mpas.process_mpas_field(h, entries, "smois", ("nSoilLevels", "nCells"), synth.mpas_field(mesh, 5, seed=3), landmask=lm)
    # masked=both is not masked=water: plain table, no fill (:1663, :1948)
    sk = np.ascontiguousarray(ts[:, 1])
    r = mpas.process_mpas_field(h, entries, "skintemp", ("nCells",), sk, landmask=lm)
    assert same(r[0]["r_arr"][None], ref_planes(0, sk))
    r = mpas.process_mpas_field(h, entries, "xice", ("nCells",), sk, landmask=lm)
    assert same(r[0]["r_arr"][None], ref_planes(0, sk, masked=True, landmask=lm, fill=0.0))
    # fields the table does not name, or that cannot be remapped, are skipped (:1648, remapper.F:339-376)
    assert mpas.process_mpas_field(h, entries, "rainnc", ("nCells",), sk) == []
    assert mpas.process_mpas_field(h, entries, "theta", ("nVertLevels", "nSomethingElse"), th) == []
    h.close()


def test_full_size_properties(orc):
    """BASELINE-size target (the 3-km CONUS mass grid of config 2, 1.9 M points) on a 40962-cell mesh, checked through properties
    that do not need the oracle at that size: every table entry is a valid node, the nearest cell is the brute-force nearest on a
    sample, weights of the cell table sum to 1, the masked table only names land cells, scaling a field by a power of two scales
    the remap exactly, a constant field comes back as the weight sum, and the fused planes equal the level-fastest remap."""
    from wps_b200 import capi
    mesh = synth.mpas_mesh(40962, seed=6)
    nx, ny, dx, tl1, tl2, slon, rlat, rlon = synth.CONFIGS["config2"]["dom"]
    h = capi.Handle(0)
    dom = synth.lambert_domain(nx, ny, dx, tl1, tl2, slon, rlat, rlon)
    xlat, xlon = h.xytoll_grid(dom, capi.M, 1, 1, nx, ny)
    d2r = np.float32(np.arcsin(np.float32(1.0)) / np.float32(90.0))
    lat, lon = np.ascontiguousarray(xlat * d2r), np.ascontiguousarray(xlon * d2r)
    h.mpas_set_mesh(mesh)
    h.mpas_remap_setup(0, lat, lon)
    near, src, w = h.mpas_get_tables(0, capi.MPAS_CELLS)
    assert src.min() >= 1 and src.max() <= mesh["nCells"] and near.min() >= 1 and near.max() <= mesh["nCells"]
    assert np.abs(w.sum(-1) - 1.0).max() < 1e-4
    rng = np.random.default_rng(0)
    pick = rng.integers(0, nx * ny, 1500)
    la, lo = lat.ravel()[pick].astype(np.float64)[:, None], lon.ravel()[pick].astype(np.float64)[:, None]
    lc, oc = mesh["latCell"].astype(np.float64)[None, :], mesh["lonCell"].astype(np.float64)[None, :]
    d = np.sin(0.5 * (lc - la)) ** 2 + np.cos(la) * np.cos(lc) * np.sin(0.5 * (oc - lo)) ** 2
    assert (d.argmin(1) + 1 == near.ravel()[pick]).mean() > 0.998           # binary32 ties aside
    _, srcm, wm = h.mpas_get_tables(0, capi.MPAS_MASKED_CELLS)
    used = wm != 0.0
    assert (mesh["landmask"][srcm[used] - 1] == 1).all() and np.abs(wm.sum(-1) - 1.0).max() < 1e-4
    f = synth.mpas_field(mesh, 8, seed=2)
    out = h.mpas_remap_field(0, capi.MPAS_CELLS, f)
    assert np.array_equal(h.mpas_remap_field(0, capi.MPAS_CELLS, f * np.float32(4.0)), out * np.float32(4.0))
    const = np.full((mesh["nCells"], 1), 3.0, np.float32)
    oc3 = h.mpas_remap_field(0, capi.MPAS_CELLS, const, nw=3)[..., 0]
    assert np.array_equal(oc3, ((np.float32(0.0) + w[..., 0] * np.float32(3.0)) + w[..., 1] * np.float32(3.0)) + w[..., 2] * np.float32(3.0))
    planes = [np.empty((ny, nx), np.float32) for _ in range(8)]
    h.mpas_remap_planes(0, capi.MPAS_CELLS, f, planes)
    assert np.array_equal(np.stack(planes), np.moveaxis(out, 2, 0))
    h.close()


def test_tiny_meshes_and_degenerate_shapes(orc):
    """The smallest meshes scipy can tessellate (12, 20, 42 cells: every walk crosses most of the sphere), a single target point, a
    single row and a single column of targets, one level with the interface rule (plane 0 only), an empty level list."""
    from wps_b200 import capi
    orc.set_transc_mode(1)
    h = capi.Handle(0)
    for ncells in (12, 20, 42):
        mesh = synth.mpas_mesh(ncells, seed=9)
        h.mpas_set_mesh(mesh)
        for shape in ((1, 1), (1, 37), (29, 1)):
            rng = np.random.default_rng(ncells + shape[0])
            lat = np.ascontiguousarray(rng.uniform(-1.5, 1.5, shape).astype(np.float32))
            lon = np.ascontiguousarray(rng.uniform(-3.1, 3.1, shape).astype(np.float32))
            t = orc.mpas_remap_setup(mesh, lat, lon)
            h.mpas_remap_setup(0, lat, lon)
            for kind, (kn, ks, kw) in KEYS.items():
                near, src, w = h.mpas_get_tables(0, kind)
                if kn:
                    assert np.array_equal(near, t[kn]), (ncells, shape, kn)
                assert np.array_equal(src, t[ks]), (ncells, shape, ks)
                assert np.array_equal(np.isfinite(w), np.isfinite(t[kw])), (ncells, shape, kw)
            _, srcc, wc = h.mpas_get_tables(0, capi.MPAS_CELLS)
            f = synth.mpas_field(mesh, 1, seed=3)
            plane = [np.empty(shape, np.float32)]
            h.mpas_remap_planes(0, capi.MPAS_CELLS, f[:, 0].copy(), plane, mode=1)
            ref = orc.mpas_extract_planes(orc.mpas_remap_field(srcc, wc, f, nw=3), 1)
            assert np.array_equal(plane[0].view(np.int32), ref[0].view(np.int32)), (ncells, shape)
    h.mpas_derive_tt([], [])
    h.close()
