"""More hand-derived known answers that pin the CPU oracle (the reference ships no tests for this path):
wind rotation, Coriolis, smoothers, land-use fractions / dominant category / landmask, categorical accumulation.
Every expected value below is worked out by hand from the cited Fortran lines, not from running any code."""
import numpy as np
import pytest

from oracle import oracle as O


def full_bits(ny, nx):
    return np.full((ny, (nx + 31) // 32), -1, np.int32)


def cgrid(nx, ny, lon_u, lon_v):
    """u (ny, nx+1) and v (ny+1, nx) longitudes of a C grid, constant per array"""
    return np.full((ny, nx + 1), lon_u, np.float32), np.full((ny + 1, nx), lon_v, np.float32)


def test_rotation_is_identity_on_the_standard_longitude():
    """rotate_winds_module.F:163-173,257,362: xlon == stdlon gives diff = 0, alpha = 0, and
    u_new = cos(0)*u + sin(0)*vbar = u, v_new = -sin(0)*ubar + cos(0)*v = v exactly, for LC and PS, both directions."""
    O.set_transc_mode(1)
    nx, ny = 7, 5
    rng = np.random.default_rng(1)
    u = rng.uniform(-30, 30, (ny, nx + 1)).astype(np.float32)
    v = rng.uniform(-30, 30, (ny + 1, nx)).astype(np.float32)
    xu, xv = cgrid(nx, ny, -98.0, -98.0)
    for code, kw in ((O.PROJ_LC, dict(truelat1=30., truelat2=60.)), (O.PROJ_PS, dict(truelat1=60.))):
        p = O.map_set(code, stdlon=-98., lat1=40., lon1=-98., knowni=4., knownj=3., dx=30000., **kw)
        for idir in (1, -1):
            uo, vo = O.metmap_xform(p, u, full_bits(ny, nx + 1), v, full_bits(ny + 1, nx), xu, xv, idir)
            assert np.array_equal(uo, u) and np.array_equal(vo, v)


@pytest.mark.parametrize("idir,eu,ev", [(1, 5.0, -3.0), (-1, -5.0, 3.0)])
def test_polar_stereographic_quarter_turn(idir, eu, ev):
    """PS, northern hemisphere: alpha = diff * rad_per_deg * hemi (:224).  90 degrees east of the standard longitude,
    u = 3 and v = 5 everywhere, every mask bit set: vbar = 4*5/4 = 5, ubar = 3, so met_to_map (idir = 1) gives
    u_new = cos(90)*3 + sin(90)*5 = 5, v_new = -sin(90)*3 + cos(90)*5 = -3; map_to_met (idir = -1) turns the other way."""
    O.set_transc_mode(1)
    nx, ny = 6, 6
    p = O.map_set(O.PROJ_PS, truelat1=60., stdlon=-100., lat1=60., lon1=-100., knowni=3., knownj=3., dx=30000.)
    assert p.hemi == 1.0
    u = np.full((ny, nx + 1), 3.0, np.float32)
    v = np.full((ny + 1, nx), 5.0, np.float32)
    xu, xv = cgrid(nx, ny, -10.0, -10.0)
    uo, vo = O.metmap_xform(p, u, full_bits(ny, nx + 1), v, full_bits(ny + 1, nx), xu, xv, idir)
    assert np.allclose(uo, eu, rtol=0, atol=1e-6) and np.allclose(vo, ev, rtol=0, atol=1e-6)


def test_lambert_rotation_uses_the_cone():
    """LC with truelat1 = truelat2 = 30: cone = sin(30 deg) = 0.5 (module_map_utils.F:1146-1153).  60 degrees east of
    stdlon: alpha = 60 * 0.5 = 30 degrees (:172).  u = 2, v = 4: u_new = cos30*2 + sin30*4 = 3.7320508,
    v_new = -sin30*2 + cos30*4 = 2.4641016."""
    O.set_transc_mode(1)
    nx, ny = 5, 4
    p = O.map_set(O.PROJ_LC, truelat1=30., truelat2=30., stdlon=-120., lat1=30., lon1=-120., knowni=3., knownj=2., dx=30000.)
    assert abs(p.cone - 0.5) < 1e-6
    u = np.full((ny, nx + 1), 2.0, np.float32)
    v = np.full((ny + 1, nx), 4.0, np.float32)
    xu, xv = cgrid(nx, ny, -60.0, -60.0)
    uo, vo = O.metmap_xform(p, u, full_bits(ny, nx + 1), v, full_bits(ny + 1, nx), xu, xv, 1)
    assert np.allclose(uo, 3.7320508, rtol=2e-6) and np.allclose(vo, 2.4641016, rtol=2e-6)


def test_rotation_averages_only_the_valid_neighbours():
    """:233-262: the V value at a U point is the mask-weighted mean of the four surrounding V points.  Interior U point
    (i, j) = (3, 2) (1-based) sees v(2,2), v(2,3), v(3,2), v(3,3); with only v(2,2) = 1 and v(3,2) = 4 valid the mean is
    2.5, so a quarter turn (PS, +90 degrees) gives u_new = 2.5.  A U point whose four V neighbours are all invalid keeps
    its value (:259-261); a U point whose own mask bit is clear is copied (:263)."""
    O.set_transc_mode(1)
    nx, ny = 5, 4
    p = O.map_set(O.PROJ_PS, truelat1=60., stdlon=-100., lat1=60., lon1=-100., knowni=3., knownj=3., dx=30000.)
    u = np.full((ny, nx + 1), 7.0, np.float32)
    v = np.zeros((ny + 1, nx), np.float32)
    v[1, 1], v[2, 1], v[1, 2], v[2, 2] = 1.0, 2.0, 4.0, 8.0     # [j-1][i-1]: v(2,2), v(2,3), v(3,2), v(3,3)
    vm = np.zeros((ny + 1, 1), np.int32)
    vm[1, 0] = (1 << 1) | (1 << 2)                                # bits of v(2,2) and v(3,2)
    um = full_bits(ny, nx + 1).copy()
    um[3, 0] &= ~(1 << 4)                                         # u(5,4) not valid
    xu, xv = cgrid(nx, ny, -10.0, -10.0)
    uo, _ = O.metmap_xform(p, u, um, v, vm, xu, xv, 1)
    assert abs(uo[1, 2] - 2.5) < 1e-6          # u(3,2)
    assert uo[3, 0] == 7.0                     # u(1,4): no valid V neighbour
    assert uo[3, 4] == 7.0                     # u(5,4): masked out, copied


def test_coriolis_parameters():
    """process_tile_module.F:1903-1904 with OMEGA_E = 7.292e-5 (constants_module.F:9): f = 2*OMEGA*sin(lat),
    e = 2*OMEGA*cos(lat)."""
    O.set_transc_mode(1)
    f, e = O.get_coriolis(np.array([[0.0, 30.0, 90.0, -30.0]], np.float32))
    assert np.allclose(f[0], [0.0, 7.292e-5, 1.4584e-4, -7.292e-5], rtol=2e-6, atol=1e-11)
    assert np.allclose(e[0, :2], [1.4584e-4, 1.4584e-4 * 0.8660254], rtol=2e-6)


def test_one_two_one_impulse_response():
    """smooth_module.F:41-55: x sweep then y sweep with weights (0.25, 0.5, 0.25): an impulse of 16 becomes
    [[1,2,1],[2,4,2],[1,2,1]] -- all intermediate values are dyadic, so the result is exact; a second pass gives the
    5x5 binomial kernel 16 * outer(b, b) / 256 with b = (1,4,6,4,1)."""
    a = np.zeros((9, 9), np.float32)
    a[4, 4] = 16.0
    s = O.smooth(a, 1, 1)
    expect = np.zeros_like(a)
    expect[3:6, 3:6] = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], np.float32)
    assert np.array_equal(s, expect)
    s2 = O.smooth(a, 2, 1)
    b = np.array([1, 4, 6, 4, 1], np.float32)
    expect2 = np.zeros_like(a)
    expect2[2:7, 2:7] = np.outer(b, b) / 16.0
    assert np.array_equal(s2, expect2)


def test_smth_desmth_special_restores_overshoots():
    """smooth_module.F:116-134: the desmoothing sweep has weights (-0.26, 1.52, -0.26), so a spike leaves negative
    side lobes; :225-234: smth-desmth_special puts back the original value wherever the result is negative and the
    original was not.  Everything else equals plain smth-desmth."""
    a = np.zeros((11, 11), np.float32)
    a[5, 5] = 100.0
    plain = O.smooth(a, 1, 2)
    special = O.smooth(a, 1, 3)
    assert (plain < 0).any()
    assert np.array_equal(special, np.where((plain < 0) & (a >= 0), a, plain))
    # the weights of either sweep pair sum to one: a constant field comes back within rounding
    c = np.full((8, 9), 250.0, np.float32)
    assert np.allclose(O.smooth(c, 1, 2), c, rtol=1e-6)


def test_fractions_dominant_category_and_landmask():
    """process_tile_module.F:1035-1051 (fractions = counts / sum, all categories = msg_fill_val when nothing was counted),
    :1123-1144 (dominant = FIRST strict maximum, msg_fill_val when no category is above zero), :570-596 (water mask:
    land = 1 when the water fraction is < 0.5, -1 when the water category holds the fill value)."""
    fillv = -999.0
    counts = np.zeros((4, 1, 4), np.float32)                 # [category][j][i], categories 11..14
    counts[:, 0, 0] = [3, 1, 0, 4]
    counts[:, 0, 1] = [2, 2, 0, 0]                           # tie: the first one wins
    counts[:, 0, 2] = [0, 0, 0, 0]                           # nothing counted
    counts[:, 0, 3] = [1, 0, 0, 1]
    fr = O.fractions(counts, fillv)
    assert np.array_equal(fr[:, 0, 0], np.array([0.375, 0.125, 0.0, 0.5], np.float32))
    assert np.array_equal(fr[:, 0, 1], np.array([0.5, 0.5, 0.0, 0.0], np.float32))
    assert np.array_equal(fr[:, 0, 2], np.full(4, fillv, np.float32))
    dom = O.dominant_category(fr, 11, fillv)
    assert list(dom[0]) == [14.0, 11.0, fillv, 11.0]
    lm = O.landmask_from_fractions(fr, 11, [14], True, fillv)   # category 14 is water
    assert list(lm[0]) == [0.0, 1.0, -1.0, 0.0]                 # 0.5 is not < 0.5 -> water


def test_accum_categorical_counts_on_aligned_grids():
    """proc_point_module.F:245-372: a source pixel is credited to the destination cell nint(rx), nint(ry) of its centre
    provided start_x <= rx <= end_x and start_y <= ry <= end_y as REALS (:253).  Destination = 1-degree lat-lon grid, cell
    (1,1) centred on (10N, 20E), cells 1..3 in both directions; source = 0.25-degree pixels whose centres sit at +-0.125
    and +-0.375 around the cell centres, so every cell has 4 x 4 pixel centres inside and none on an edge.  The real
    compare drops pixels with rx < 1 or rx > 3 (likewise y), i.e. the outer half of the border cells: corner cells keep
    2 x 2 pixels, edge cells 2 x 4, the centre cell all 16.  Category = 1 + (pixel column mod 2) splits each count evenly."""
    O.set_transc_mode(1)
    dom = O.map_set(O.PROJ_LATLON, latinc=1.0, loninc=1.0, lat1=10.0, lon1=20.0, knowni=1.0, knownj=1.0, stdlon=0.0)
    src = O.push_source_projection(O.PROJ_LATLON, dlat=0.25, dlon=0.25, known_x=1.0, known_y=1.0, known_lat=9.625,
                                   known_lon=19.625)
    nx = ny = 3
    tn = 12
    tile = (1.0 + (np.arange(tn)[None, :] % 2) + np.zeros((tn, 1))).astype(np.float32)
    dst = np.full((2, ny, nx), 77.0, np.float32)  # "uninitialised": first touch must zero it (:330-334)
    processed = np.zeros((ny, 1), np.int32)
    new_pts = np.zeros((ny, 1), np.int32)
    ninv = O.accum_categorical(src, dom, O.M, tile, 1, 1, 0, dst, 1, 1, 1, processed, new_pts, 1, 0.0)
    assert ninv == 0
    per_cat = np.array([[2, 4, 2], [4, 8, 4], [2, 4, 2]], np.float32)
    assert np.array_equal(dst, np.stack([per_cat, per_cat]))
    assert all(int(w) & 0b111 == 0b111 for w in new_pts[:, 0])
    # a category outside [start_z, end_z] is reported, not counted (:343-351); the cell is still zeroed at first touch
    # (:330-334).  The 16 pixels of the centre cell carry category 9 here.
    tile2 = tile.copy()
    tile2[4:8, 4:8] = 9.0                         # the 16 pixels of the centre cell
    dst2 = np.full((2, ny, nx), 77.0, np.float32)
    processed[:] = 0
    new_pts[:] = 0
    ninv2 = O.accum_categorical(src, dom, O.M, tile2, 1, 1, 0, dst2, 1, 1, 1, processed, new_pts, 1, 0.0)
    assert ninv2 == 16
    assert dst2[0, 1, 1] == 0.0 and dst2[1, 1, 1] == 0.0 and dst2[0, 0, 0] == 2.0


def test_geogrid_golden_regression():
    """the oracle still produces the committed geogrid fixtures (tests/golden/make_golden_geogrid.py)"""
    import os
    from tests.golden import make_golden_geogrid as mg
    g = np.load(os.path.join(os.path.dirname(mg.__file__), "geogrid_small.npz"))
    now = mg.build()
    assert sorted(g.files) == sorted(now.keys())
    for k in g.files:
        a, b = g[k], now[k]
        assert a.shape == b.shape, k
        assert np.array_equal(a.view(np.int32) if a.dtype == np.float32 else a, b.view(np.int32) if b.dtype == np.float32 else b), k


def test_rotated_latlon_known_answers(orc):
    """PROJ_ROTLL (module_map_utils.F:1659-1909), hand-derived: the grid's centre (lat1, lon1) is rotated (0, 0), i.e. column
    ixdim of the 2*ixdim-1 interleaved columns and row jydim/2+1; on an odd row a mass point sits at col/2 + 0.5 and a velocity
    point at col/2 (:1731-1737, :1773-1779); along the central meridian the rotated latitude equals lat - lat1, so the row moves
    by 1 per dphd = phi / ((jydim-1)/2) degrees; xytoll inverts lltoxy on grid points to the 0.0002 nudge of :1723-1727."""
    orc.set_transc_mode(1)
    kw = dict(ixdim=120, jydim=241, phi=12.0, lat1=38.0, lon1=-97.0)
    kw["lambda"] = 10.0
    f = np.float32
    p = orc.map_set(orc.PROJ_ROTLL, stagger=orc.HH, **kw)
    x, y = orc.lltoxy(p, np.array([38.0], f), np.array([-97.0], f), stagger=orc.HH)
    assert x[0] == 60.5 and y[0] == 121.0
    x, y = orc.lltoxy(p, np.array([38.0], f), np.array([-97.0], f), stagger=orc.VV)
    assert x[0] == 60.0 and y[0] == 121.0
    dphd = 12.0 / 120.0
    x, y = orc.lltoxy(p, np.array([38.0 + 10 * dphd], f), np.array([-97.0], f), stagger=orc.HH)
    assert abs(y[0] - 131.0) < 2e-4 + 1e-4 and abs(x[0] - 60.5) < 1e-4           # row 131 is odd as well
    x, y = orc.lltoxy(p, np.array([38.0 + 11 * dphd], f), np.array([-97.0], f), stagger=orc.HH)
    assert abs(y[0] - 132.0) < 3e-4 and abs(x[0] - 60.0) < 1e-4                   # even row: no half-column shift
    rng = np.random.default_rng(3)
    for st in (orc.HH, orc.VV):
        ii = rng.integers(1, 120, 500).astype(f)
        jj = rng.integers(1, 241, 500).astype(f)
        la, lo = orc.xytoll(p, ii, jj, stagger=st)
        x2, y2 = orc.lltoxy(p, la, lo, stagger=st)
        assert np.abs(x2 - ii).max() < 3e-4 and np.abs(y2 - jj).max() < 3e-4
    # a velocity point with the indices of a mass point lies one column of the interleaved grid (dlmd) to its east on odd rows,
    # to its west on even rows (:1868-1874)
    p.stagger = orc.HH
    la_h, lo_h = orc.xytoll(p, np.array([60.0, 60.0], f), np.array([121.0, 122.0], f), stagger=orc.HH)
    la_v, lo_v = orc.xytoll(p, np.array([60.0, 60.0], f), np.array([121.0, 122.0], f), stagger=orc.VV)
    assert lo_v[0] > lo_h[0] and lo_v[1] < lo_h[1]
    # map_set demands ixdim, jydim, phi, lambda, lat1, lon1, stagger (:394-405)
    with pytest.raises(ValueError):
        orc.map_set(orc.PROJ_ROTLL, ixdim=120, jydim=241, phi=12.0, lat1=38.0, lon1=-97.0, stagger=orc.HH)


def test_nmm_wind_rotation_known_answers(orc):
    """metmap_xform_nmm (rotate_winds_module.F:455-631): on the central meridian of the rotated grid the angle is zero; elsewhere
    it is a pure rotation (speed kept, idir = +1 then -1 is the identity); a masked-out partner counts as 0 and a masked-out
    component is copied (:519-541)."""
    orc.set_transc_mode(1)
    f = np.float32
    p = orc.map_set(orc.PROJ_ROTLL, ixdim=40, jydim=61, phi=6.0, lat1=40.0, lon1=-100.0, stagger=orc.VV, **{"lambda": 8.0})
    ny, nx = 5, 7
    xlat = (40.0 + np.arange(ny, dtype=f)[:, None] * 2.0 + np.zeros((1, nx), f)).astype(f)
    xlon = (-100.0 + (np.arange(nx, dtype=f)[None, :] - 3.0) * 4.0 + np.zeros((ny, 1), f)).astype(f)
    rng = np.random.default_rng(5)
    u, v = rng.uniform(-20, 20, (ny, nx)).astype(f), rng.uniform(-20, 20, (ny, nx)).astype(f)
    full = np.full((ny, 1), -1, np.int32)
    u1, v1 = orc.metmap_xform_nmm(p, u, full, v, full, xlat, xlon, 1)
    # central meridian: alpha = 0 up to the rounding of relm = -100 deg - 260 deg = -2 pi (lon1 < 0 is moved to 0..360, :483-487)
    assert np.abs(u1[:, 3] - u[:, 3]).max() < 1e-5 and np.abs(v1[:, 3] - v[:, 3]).max() < 1e-5
    assert np.abs(np.hypot(u1, v1) - np.hypot(u, v)).max() < 1e-4
    assert np.abs(u1[:, 0] - u[:, 0]).max() > 0.1                                            # off it the wind does turn
    u2, v2 = orc.metmap_xform_nmm(p, u1, full, v1, full, xlat, xlon, -1)
    assert np.abs(u2 - u).max() < 1e-4 and np.abs(v2 - v).max() < 1e-4
    # masks: bit 0 of each row cleared in v_mask -> U there is cos(alpha) * u, V is copied
    vm = np.full((ny, 1), -2, np.int32)
    u3, v3 = orc.metmap_xform_nmm(p, u, full, v, vm, xlat, xlon, 1)
    assert np.array_equal(v3[:, 0], v[:, 0]) and np.array_equal(u3[:, 1:], u1[:, 1:]) and np.all(np.abs(u3[:, 0]) <= np.abs(u[:, 0]))
    # the Lambert branch of the same routine (:560-612) equals the C-grid rotation at collocated points with full masks
    pl = orc.map_set(orc.PROJ_LC, truelat1=30., truelat2=60., stdlon=-98., lat1=34.83, lon1=-81.03, knowni=30., knownj=25., dx=30000.)
    u4, v4 = orc.metmap_xform_nmm(pl, u, full, v, full, xlat, xlon, -1)
    diff = -1.0 * (xlon.astype(np.float64) + 98.0)
    alpha = np.deg2rad(diff) * float(pl.cone)
    assert np.abs(u4 - (np.cos(alpha) * u - np.sin(alpha) * v)).max() < 1e-4
    assert np.abs(v4 - (np.sin(alpha) * u + np.cos(alpha) * v)).max() < 1e-4


def _py_egrid_smooth(a, npass, opt, msgval, hflag, sx, sy):
    """one_two_one_egrid / smth_desmth_egrid written from smooth_module.F:252-338, :448-585 with Python loops (one level)."""
    f = np.float32
    a = a.copy()
    ny, nx = a.shape
    ex, ey = sx + nx - 1, sy + ny - 1
    A = lambda arr, ix, iy: arr[iy - sy, ix - sx]          # noqa: E731
    ihe, ihw, istart = {}, {}, {}
    for iy in range(sy, ey + 1):
        m = lambda v: int(np.fmod(v, 2))                   # noqa: E731  Fortran MOD truncates
        ihe[iy] = abs(m(iy + 1)) if hflag == 1.0 else abs(m(iy))
        ihw[iy] = ihe[iy] - 1
        istart[iy] = (sx if m(iy) == 0 else sx + 1) if hflag == 1.0 else (sx if abs(m(iy)) == 1 else sx + 1)
    cond = lambda ix, iy: (msgval == 1.0 and A(a, ix, iy) != 0.0) or msgval != 1.0     # noqa: E731
    for _ in range(npass):
        s = a.copy()
        for cw, ew, lo in ((f(0.50), f(0.25), 1),) + (((f(1.52), f(-0.26), 2),) if opt == 2 else ()):
            for iy in range(sy + lo, ey - lo + 1):
                for ix in range(istart[iy], ex):
                    if cond(ix, iy):
                        s[iy - sy, ix - sx] = f(f(cw * A(a, ix, iy)) + f(ew * f(A(a, ix + ihw[iy], iy - 1) + A(a, ix + ihe[iy], iy + 1))))
            for iy in range(sy + lo, ey - lo + 1):
                for ix in range(istart[iy], ex):
                    if cond(ix, iy):
                        a[iy - sy, ix - sx] = f(f(cw * A(s, ix, iy)) + f(ew * f(A(s, ix + ihe[iy], iy - 1) + A(s, ix + ihw[iy], iy + 1))))
    return a


@pytest.mark.parametrize("opt", [1, 2])
@pytest.mark.parametrize("hflag", [1.0, 0.0])
def test_egrid_smoothers(orc, opt, hflag):
    rng = np.random.default_rng(int(opt * 2 + hflag))
    a = rng.uniform(-5, 50, (1, 14, 11)).astype(np.float32)
    a[0, 4:7, 3:6] = 0.0
    for msgval, sx, sy in ((1.0e20, 1, 1), (1.0, 1, 1), (1.0e20, -2, -2), (1.0, 0, 3)):
        for npass in (1, 2):
            got = orc.smooth_egrid(a, npass, opt, msgval=msgval, hflag=hflag, sx=sx, sy=sy)
            ref = _py_egrid_smooth(a[0], npass, opt, msgval, hflag, sx, sy)
            assert np.array_equal(got[0].view(np.int32), ref.view(np.int32)), (msgval, sx, sy, npass)
            if msgval == 1.0:
                assert (got[0, 4:7, 3:6] == 0.0).all()                       # zeros are left alone (:296)
    c = np.full((1, 9, 9), 7.0, np.float32)
    assert np.array_equal(orc.smooth_egrid(c, 3, 1, hflag=hflag), c)          # 0.5 c + 0.25 (2 c) = c
    # smth-desmth_special: 'not currently implemented for NMM' (process_tile_module.F:675)
    assert np.array_equal(orc.smooth_egrid(a, 2, 3, hflag=hflag), a)
