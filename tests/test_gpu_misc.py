"""GPU parity: projections, wind rotation, smoothers, fractions/dominant/landmask, geometry outputs."""
import numpy as np
import pytest

from tests.util import assert_bitexact, bool_to_bits, conus_like, ulp_diff

pytestmark = pytest.mark.gpu


def proj_cases(orc, capi):
    cases = {
        "lc": (orc.PROJ_LC, dict(truelat1=30., truelat2=60., stdlon=-98., lat1=34.83, lon1=-81.03, knowni=50.5, knownj=50.5, dx=30000.)),
        "lc_tangent": (orc.PROJ_LC, dict(truelat1=38.5, truelat2=38.5, stdlon=-97.5, lat1=38.5, lon1=-97.5, knowni=900., knownj=530., dx=3000.)),
        "lc_south": (orc.PROJ_LC, dict(truelat1=-30., truelat2=-60., stdlon=140., lat1=-35., lon1=145., knowni=40., knownj=40., dx=12000.)),
        "ps": (orc.PROJ_PS, dict(truelat1=60., stdlon=-100., lat1=70., lon1=-110., knowni=60., knownj=60., dx=25000.)),
        "ps_south": (orc.PROJ_PS, dict(truelat1=-71., stdlon=0., lat1=-75., lon1=20., knowni=50., knownj=50., dx=30000.)),
        "merc": (orc.PROJ_MERC, dict(truelat1=10., lat1=5., lon1=100., knowni=30., knownj=30., dx=27000.)),
        "latlon": (orc.PROJ_LATLON, dict(latinc=-0.25, loninc=0.25, knowni=1., knownj=1., lat1=90., lon1=0., nxmax=1440)),
        "cassini": (orc.PROJ_CASSINI, dict(latinc=0.5, loninc=0.5, lat1=10., lon1=-20., lat0=60., lon0=-140., knowni=1., knownj=1., stdlon=-140., dx=55000., dy=55000.)),
        "cyl": (orc.PROJ_CYL, dict(latinc=1.0, loninc=1.0, stdlon=0.)),
        "ps_wgs84": (orc.PROJ_PS_WGS84, dict(truelat1=70., stdlon=-45., lat1=65., lon1=-50., knowni=100., knownj=100., dx=5000.)),
        "albers": (orc.PROJ_ALBERS_NAD83, dict(truelat1=29.5, truelat2=45.5, stdlon=-96., lat1=23., lon1=-96., knowni=1., knownj=1., dx=1000.)),
    }
    return cases


def test_lltoxy_xytoll_all_projections(orc, handle):
    """(lat,lon)<->(i,j) for every projection: bit-exact vs the oracle under the shared FP64-then-round policy
    (<=2e-6 of the points may differ by 1 ulp from binary64 libm differences); ulp histogram vs glibc float libm reported."""
    from wps_b200 import capi
    rng = np.random.default_rng(0)
    n = 200000
    for name, (code, kw) in proj_cases(orc, capi).items():
        if name == "cyl":
            continue  # set_cyl domains carry no knowni/lat1; exercised through cassini
        orc.set_transc_mode(1)
        po = orc.map_set(code, **kw)
        pg = capi.map_set(code, **kw)
        for f in ("cone", "polei", "polej", "rsw", "rebydx", "hemi", "dlon", "lat1", "lon1", "rho0", "nc", "bigc"):
            assert np.float32(getattr(po, f)).tobytes() == np.float32(getattr(pg, f)).tobytes(), (name, f)
        x = rng.uniform(-20, 220, n).astype(np.float32)
        y = rng.uniform(-20, 220, n).astype(np.float32)
        for stag in (capi.M, capi.U, capi.V, capi.CORNER):
            lat_g, lon_g = handle.xytoll(pg, x, y, stag)
            lat_o, lon_o = orc.xytoll(po, x, y, stag)
            ok = np.isfinite(lat_o) & np.isfinite(lon_o)
            d = np.maximum(ulp_diff(lat_g[ok], lat_o[ok]), ulp_diff(lon_g[ok], lon_o[ok]))
            assert (d > 1).mean() < 2e-5 and (d > 0).mean() < 2e-5, (name, "xytoll", int((d > 0).sum()), int(d.max()))
            # forward on the oracle's own lat/lon so both sides see identical inputs
            la = np.where(ok, lat_o, 0).astype(np.float32)
            lo = np.where(ok, lon_o, 0).astype(np.float32)
            xg, yg = handle.lltoxy(pg, la, lo, stag)
            xo, yo = orc.lltoxy(po, la, lo, stag)
            d = np.maximum(ulp_diff(xg, xo), ulp_diff(yg, yo))
            # a 1-ulp input difference can flip a periodic-wrap branch (Cassini/cylindrical), hence count-based
            assert (d > 1).mean() < 2e-5 and (d > 0).mean() < 2e-5, (name, "lltoxy", int((d > 0).sum()), int(d.max()))
        orc.set_transc_mode(0)
        xo0, yo0 = orc.lltoxy(po, la, lo, capi.M)
        d0 = np.maximum(ulp_diff(xg if stag == capi.M else handle.lltoxy(pg, la, lo, capi.M)[0], xo0), 0)
        print("proj %-10s vs glibc-float libm: %.4f%% of x differ, max %d ulp" % (name, 100 * (d0 > 0).mean(), d0.max()))
    orc.set_transc_mode(1)


def test_rotate_winds(orc, handle):
    from wps_b200 import capi
    rng = np.random.default_rng(1)
    for code, kw in ((orc.PROJ_LC, dict(truelat1=30., truelat2=60., stdlon=-98., lat1=34.83, lon1=-81.03, knowni=30., knownj=25., dx=30000.)),
                     (orc.PROJ_PS, dict(truelat1=60., stdlon=-100., lat1=70., lon1=-110., knowni=30., knownj=25., dx=25000.))):
        orc.set_transc_mode(1)
        po, pg = orc.map_set(code, **kw), capi.map_set(code, **kw)
        nx, ny, nlev = 61, 49, 4
        _, xlon_u = orc.get_lat_lon_fields(po, 1, 1, nx + 1, ny, orc.U)
        _, xlon_v = orc.get_lat_lon_fields(po, 1, 1, nx, ny + 1, orc.V)
        u = rng.uniform(-30, 30, (nlev, ny, nx + 1)).astype(np.float32)
        v = rng.uniform(-30, 30, (nlev, ny + 1, nx)).astype(np.float32)
        um = np.stack([bool_to_bits(rng.random((ny, nx + 1)) < p) for p in (1.0, 0.9, 0.5, 0.0)])
        vm = np.stack([bool_to_bits(rng.random((ny + 1, nx)) < p) for p in (1.0, 0.8, 0.5, 1.0)])
        for idir in (1, -1):
            ug, vg = u.copy(), v.copy()
            handle.rotate_winds(idir, pg, ug, um, vg, vm, xlon_u, xlon_v)
            for l in range(nlev):
                uo, vo = orc.metmap_xform(po, u[l], um[l], v[l], vm[l], xlon_u, xlon_v, idir)
                d = max(ulp_diff(ug[l], uo).max(), ulp_diff(vg[l], vo).max())
                frac = max((ulp_diff(ug[l], uo) > 0).mean(), (ulp_diff(vg[l], vo) > 0).mean())
                # sin/cos of alpha are FP64-then-round on both sides: identical except rare double-rounding cases
                assert d <= 1 and frac < 1e-4, (code, idir, l, d, frac)


@pytest.mark.parametrize("opt", [1, 2, 3])
@pytest.mark.parametrize("npass", [1, 3])
def test_smoothers(orc, handle, opt, npass):
    rng = np.random.default_rng(opt * 10 + npass)
    a = rng.uniform(-50, 3000, (3, 67, 85)).astype(np.float32)
    a[0, 10:20, 10:30] = 0.0
    g = a.copy()
    handle.smooth(g, opt, npass)
    assert_bitexact(g, orc.smooth(a, npass, opt), "smooth opt=%d npass=%d" % (opt, npass))


def test_fractions_dominant_landmask(orc, handle):
    rng = np.random.default_rng(4)
    ncat, ny, nx = 21, 40, 50
    counts = rng.integers(0, 6, (ncat, ny, nx)).astype(np.float32)
    counts[:, :3, :] = 0.0  # cells with no data -> msg_fill_val
    counts[16, 10:20, :] += 40  # water (category 17)
    for is_water, lm_vals in ((True, [17, 21]), (False, [1, 2, 3])):
        f = counts.copy()
        lm = np.empty((ny, nx), np.float32)
        dom = np.empty((ny, nx), np.float32)
        handle.fractions_dominant(f, 1, 1e20, lm_vals, is_water, lm, dom)
        fo = orc.fractions(counts, 1e20)
        assert_bitexact(f, fo, "fractions")
        lmo = orc.landmask_from_fractions(fo, 1, lm_vals, is_water, 1e20)
        assert_bitexact(lm, lmo, "landmask")
        assert_bitexact(dom, orc.dominant_category_landmask(fo, lmo, 1, lm_vals, is_water), "dominant(landmask field)")
    f = counts.copy()
    dom = np.empty((ny, nx), np.float32)
    handle.fractions_dominant(f, 1, 1e20, (), True, None, dom)
    assert_bitexact(dom, orc.dominant_category(orc.fractions(counts, 1e20), 1, 1e20), "dominant")


def test_geometry_outputs(orc, handle):
    from wps_b200 import capi
    orc.set_transc_mode(1)
    kw = dict(truelat1=30., truelat2=60., stdlon=-98., lat1=34.83, lon1=-81.03, knowni=50.5, knownj=50.5, dx=30000.)
    po, pg = orc.map_set(orc.PROJ_LC, **kw), capi.map_set(capi.PROJ_LC, **kw)
    for stag, nx, ny in ((capi.M, 100, 100), (capi.U, 101, 100), (capi.V, 100, 101), (capi.CORNER, 101, 101)):
        lat_g, lon_g = handle.xytoll_grid(pg, stag, -2, -2, nx + 6, ny + 6)
        lat_o, lon_o = orc.get_lat_lon_fields(po, -2, -2, nx + 6, ny + 6, stag)
        d = np.maximum(ulp_diff(lat_g, lat_o), ulp_diff(lon_g, lon_o))
        assert d.max() <= 1 and (d > 0).mean() < 1e-4
    lat_o, lon_o = orc.get_lat_lon_fields(po, 1, 1, 100, 100, orc.M)
    for tl2 in (60., 30.):
        mxg, myg = handle.map_factor(capi.PROJ_LC, 30., tl2, lat_o, lon_o)
        mxo, myo = orc.get_map_factor(po, lat_o, lon_o, 30., tl2)
        assert ulp_diff(mxg, mxo).max() <= 1 and ulp_diff(myg, myo).max() <= 1
    fg, eg = handle.coriolis(lat_o)
    fo, eo = orc.get_coriolis(lat_o)
    assert ulp_diff(fg, fo).max() <= 1 and ulp_diff(eg, eo).max() <= 1
    cg, sg = handle.rotang(capi.PROJ_LC, 30., 60., -98., lat_o, lon_o)
    co, so = orc.get_rotang(orc.PROJ_LC, lat_o, lon_o, 30., 60., -98.)
    assert ulp_diff(cg, co).max() <= 1 and ulp_diff(sg, so).max() <= 1
    rng = np.random.default_rng(8)
    hgt = rng.uniform(0, 3000, (2, 100, 100)).astype(np.float32)
    mf = rng.uniform(0.9, 1.1, (100, 100)).astype(np.float32)
    for m in (mf, None):
        assert_bitexact(handle.dfdx(hgt, 1, 30000., m), orc.calc_dfd("x", hgt, 1, m, 30000.), "dfdx")
        assert_bitexact(handle.dfdy(hgt, 1, 30000., m), orc.calc_dfd("y", hgt, 1, m, 30000.), "dfdy")


@pytest.mark.parametrize("idir", [1, -1])
def test_rotate_winds_cassini(orc, handle, idir):
    """PROJ_CASSINI branches of metmap_xform (rotate_winds_module.F:174-222, 281-329): the rotation angle comes from
    finite differences (of ij_to_latlon at y +- 0.1 for map_to_met, of the domain's lat/lon arrays along j for
    met_to_map) instead of the standard-longitude formula; needs xlat_u / xlat_v besides xlon."""
    from wps_b200 import capi
    rng = np.random.default_rng(3)
    kw = dict(latinc=0.5, loninc=0.5, lat1=10., lon1=-20., lat0=60., lon0=-140., knowni=1., knownj=1., stdlon=-140., dx=55000., dy=55000.)
    orc.set_transc_mode(1)
    po, pg = orc.map_set(orc.PROJ_CASSINI, **kw), capi.map_set(capi.PROJ_CASSINI, **kw)
    nx, ny, nlev = 47, 39, 3
    xlat_u, xlon_u = orc.get_lat_lon_fields(po, 1, 1, nx + 1, ny, orc.U)
    xlat_v, xlon_v = orc.get_lat_lon_fields(po, 1, 1, nx, ny + 1, orc.V)
    u = rng.uniform(-30, 30, (nlev, ny, nx + 1)).astype(np.float32)
    v = rng.uniform(-30, 30, (nlev, ny + 1, nx)).astype(np.float32)
    um = np.stack([bool_to_bits(rng.random((ny, nx + 1)) < p) for p in (1.0, 0.8, 0.5)])
    vm = np.stack([bool_to_bits(rng.random((ny + 1, nx)) < p) for p in (1.0, 0.9, 0.5)])
    ug, vg = u.copy(), v.copy()
    handle.rotate_winds(idir, pg, ug, um, vg, vm, xlon_u, xlon_v, xlat_u=xlat_u, xlat_v=xlat_v)
    changed = 0
    for l in range(nlev):
        uo, vo = orc.metmap_xform(po, u[l], um[l], v[l], vm[l], xlon_u, xlon_v, idir, xlat_u=xlat_u, xlat_v=xlat_v)
        d = max(ulp_diff(ug[l], uo).max(), ulp_diff(vg[l], vo).max())
        frac = max((ulp_diff(ug[l], uo) > 0).mean(), (ulp_diff(vg[l], vo) > 0).mean())
        # every transcendental is FP64-then-round on both sides; the rotated-pole transform feeding alpha (idir=+1) may
        # flip a last bit in rare double-rounding cases, which cos/sin(alpha) carry into the result
        assert d <= 2 and frac < 2e-3, (idir, l, d, frac)
        changed += int((uo != u[l]).sum())
    assert changed > 0
    # without the latitudes the Cassini rotation is refused loudly
    with pytest.raises(capi.WpsGpuError):
        handle.rotate_winds(idir, pg, u.copy(), um, v.copy(), vm, xlon_u, xlon_v)


def test_rotated_latlon_nmm_egrid(orc, handle):
    """PROJ_ROTLL (NMM E-grid, module_map_utils.F:1659-1909) with the HH / VV staggers lltoxy and xytoll switch the projection
    to (llxy_module.F:800-804, :851-862): device against the oracle on the same inputs -- binary64 arithmetic on both sides,
    identical except where CUDA's and glibc's binary64 functions round differently -- and the E-grid geometry outputs."""
    from wps_b200 import capi
    rng = np.random.default_rng(11)
    n = 200000
    kw = dict(ixdim=120, jydim=241, phi=12.0, lat1=38.0, lon1=-97.0)
    kw["lambda"] = 10.0
    orc.set_transc_mode(1)
    for map_stag in (capi.HH, capi.VV):
        po, pg = orc.map_set(orc.PROJ_ROTLL, stagger=map_stag, **kw), capi.map_set(capi.PROJ_ROTLL, stagger=map_stag, **kw)
        for stag in (capi.HH, capi.VV, capi.M):      # M: the stagger map_set was given stays in force
            x = rng.uniform(-5, 130, n).astype(np.float32)
            y = rng.uniform(-5, 250, n).astype(np.float32)
            x[:2000] = np.round(x[:2000]); y[:2000] = np.round(y[:2000])       # grid points: the 0.999 / 0.0002 nudges
            po.stagger = map_stag
            lat_o, lon_o = orc.xytoll(po, x, y, stag)
            lat_g, lon_g = handle.xytoll(pg, x, y, stag)
            d = np.maximum(ulp_diff(lat_g, lat_o), ulp_diff(lon_g, lon_o))
            assert (d > 0).mean() < 1e-4 and d.max() <= 2, ("xytoll", map_stag, stag, (d > 0).mean(), d.max())
            po.stagger = map_stag
            xo, yo = orc.lltoxy(po, lat_o, lon_o, stag)
            xg, yg = handle.lltoxy(pg, lat_o, lon_o, stag)
            d = np.maximum(ulp_diff(xg, xo), ulp_diff(yg, yo))
            assert (d > 0).mean() < 1e-4 and d.max() <= 2, ("lltoxy", map_stag, stag, (d > 0).mean(), d.max())
    # get_lat_lon_fields on the E-grid: mass and velocity points
    po, pg = orc.map_set(orc.PROJ_ROTLL, stagger=capi.HH, **kw), capi.map_set(capi.PROJ_ROTLL, stagger=capi.HH, **kw)
    for stag in (capi.HH, capi.VV):
        po.stagger = capi.HH
        la_o, lo_o = orc.get_lat_lon_fields(po, 1, 1, 119, 241, stag)
        la_g, lo_g = handle.xytoll_grid(pg, stag, 1, 1, 119, 241)
        d = np.maximum(ulp_diff(la_g, la_o), ulp_diff(lo_g, lo_o))
        assert (d > 0).mean() < 5e-4 and d.max() <= 2, (stag, (d > 0).mean(), d.max())      # 28 k points: a handful of 1-ulp cases


def test_rotate_winds_nmm(orc, handle):
    """metmap_xform_nmm on the device (E-grid, U and V collocated), both directions, rotated lat-lon and Lambert branches,
    partial masks, several levels in one call: <= 1 ulp against the oracle (sin / cos / asin FP64-then-round on both sides)."""
    from wps_b200 import capi
    rng = np.random.default_rng(21)
    orc.set_transc_mode(1)
    kw = dict(ixdim=60, jydim=91, phi=9.0, lat1=38.0, lon1=-97.0, stagger=capi.VV)
    kw["lambda"] = 12.0
    cases = [(orc.map_set(orc.PROJ_ROTLL, **kw), capi.map_set(capi.PROJ_ROTLL, **kw))]
    lc = dict(truelat1=30., truelat2=60., stdlon=-98., lat1=34.83, lon1=-81.03, knowni=30., knownj=25., dx=30000.)
    cases.append((orc.map_set(orc.PROJ_LC, **lc), capi.map_set(capi.PROJ_LC, **lc)))
    po0 = cases[0][0]
    xlat, xlon = orc.get_lat_lon_fields(po0, 1, 1, 59, 91, orc.VV)
    ny, nx = xlat.shape
    nlev = 4
    u = rng.uniform(-30, 30, (nlev, ny, nx)).astype(np.float32)
    v = rng.uniform(-30, 30, (nlev, ny, nx)).astype(np.float32)
    um = np.stack([bool_to_bits(rng.random((ny, nx)) < p) for p in (1.0, 0.9, 0.5, 0.0)])
    vm = np.stack([bool_to_bits(rng.random((ny, nx)) < p) for p in (1.0, 0.8, 0.5, 1.0)])
    for po, pg in cases:
        for idir in (1, -1):
            ug, vg = u.copy(), v.copy()
            handle.rotate_winds_nmm(idir, pg, ug, um, vg, vm, xlat, xlon)
            for l in range(nlev):
                uo, vo = orc.metmap_xform_nmm(po, u[l], um[l], v[l], vm[l], xlat, xlon, idir)
                d = max(ulp_diff(ug[l], uo).max(), ulp_diff(vg[l], vo).max())
                frac = max((ulp_diff(ug[l], uo) > 0).mean(), (ulp_diff(vg[l], vo) > 0).mean())
                assert d <= 2 and frac < 1e-3, (po.code, idir, l, d, frac)
    # a projection the routine does not rotate for leaves the arrays alone (:473, :560)
    pm = capi.map_set(capi.PROJ_MERC, truelat1=10., lat1=5., lon1=100., knowni=30., knownj=30., dx=27000.)
    ug, vg = u.copy(), v.copy()
    handle.rotate_winds_nmm(1, pm, ug, um, vg, vm, xlat, xlon)
    assert np.array_equal(ug, u) and np.array_equal(vg, v)


@pytest.mark.parametrize("opt", [1, 2, 3])
@pytest.mark.parametrize("hflag", [1.0, 0.0])
def test_egrid_smoothers_on_device(orc, handle, opt, hflag):
    """one_two_one_egrid / smth_desmth_egrid through the C ABI, bit for bit: several levels, the msgval == 1.0 rule that leaves
    zeros alone, odd and even lower bounds (the row parity decides which diagonal neighbours a point has)."""
    rng = np.random.default_rng(int(10 * opt + hflag))
    a = rng.uniform(-50, 3000, (3, 67, 85)).astype(np.float32)
    a[:, 10:20, 10:30] = 0.0
    for msgval, sx, sy, npass in ((1.0e20, 1, 1, 1), (1.0, 1, 1, 2), (1.0e20, -2, -2, 3), (1.0, 4, 7, 1)):
        g = a.copy()
        handle.smooth_egrid(g, opt, npass, msgval=msgval, hflag=hflag, sx=sx, sy=sy)
        assert_bitexact(g, orc.smooth_egrid(a, npass, opt, msgval=msgval, hflag=hflag, sx=sx, sy=sy), "egrid smooth opt=%d" % opt)
