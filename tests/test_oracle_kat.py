"""CPU: hand-derived known-answer tests for the oracle, each derived from the Fortran source line cited.
These (plus tests/test_oracle_crosscheck.py and the golden regression) are the manufactured pins of an oracle the
reference itself cannot pin (it ships no tests)."""
import numpy as np
import pytest

F = np.float32


def grid(nx=12, ny=10):
    j, i = np.mgrid[1:ny + 1, 1:nx + 1]
    return (i + 10 * j).astype(np.float32)


def test_interp_option_parsing(orc):
    """interp_module.F:20-295, misc_definitions_module.F:21-23"""
    assert orc.parse_interp_option("sixteen_pt+four_pt+wt_average_4pt+wt_average_16pt+search") == ([1, 2, 6, 7, 8, 0], [0, 0, 0, 0, 1200, 0])
    assert orc.parse_interp_option("four_pt+average_4pt") == ([2, 4, 0], [0, 0, 0])
    assert orc.parse_interp_option("average_16pt+nearest_neighbor") == ([5, 3, 0], [0, 0, 0])
    # average_gcell tokens are skipped, not listed (:81, :122); threshold parsed separately
    assert orc.parse_interp_option("average_gcell(4.0)+four_pt+average_4pt") == ([2, 4, 0, 0], [0, 0, 0, 0])
    assert orc.gcell_threshold("average_gcell(4.0)+four_pt+average_4pt") == 4.0
    # search depth in parentheses (:254-295); default 1200
    assert orc.parse_interp_option("four_pt+search(5)") == ([2, 8, 0], [0, 5, 0])
    assert orc.parse_interp_option("search") == ([8, 0], [1200, 0])
    # the trailing token is parsed only `if (p1 < iend)` (:89): a one-character tail is dropped silently
    assert orc.parse_interp_option("four_pt+x") == ([2, 0, 0], [0, 0, 0])
    # exact-length match: wt_average_4pt is not average_4pt (:53-64)
    assert orc.parse_interp_option("wt_average_4pt")[0] == [6, 0]


def test_nearest_neighbor_rounds_half_away_and_stops_outside(orc):
    a = grid()
    # nint(2.5) = 3 (:402-403)
    assert orc.interp_sequence([2.5], [2.5], a, 1, 1, 1e20, "nearest_neighbor")[0] == 3 + 30
    assert orc.interp_sequence([2.49], [3.5], a, 1, 1, 1e20, "nearest_neighbor")[0] == 2 + 40
    # outside the array: msgval is returned WITHOUT trying the next method (:406-414)
    assert orc.interp_sequence([0.4], [3.0], a, 1, 1, 1e20, "nearest_neighbor+search")[0] == F(1e20)
    # a missing value does go on to the next method (:437-440)
    b = a.copy()
    b[2, 2] = -1e30
    assert orc.interp_sequence([3.0], [3.0], b, 1, 1, -1e30, "nearest_neighbor+search")[0] != F(-1e30)


def test_four_pt_is_exact_on_bilinear_field(orc):
    """:1165-1168: a(i,j)=i+10j is reproduced exactly at dyadic fractions"""
    a = grid()
    xx = np.array([2.5, 3.25, 7.75, 4.0, 5.5], np.float32)
    yy = np.array([2.5, 4.75, 1.125, 6.5, 3.0], np.float32)
    got = orc.interp_sequence(xx, yy, a, 1, 1, 1e20, "four_pt")
    assert np.array_equal(got, xx + 10 * yy)
    # any corner outside -> next method; none left -> msgval (:1086-1096, :328-331)
    assert orc.interp_sequence([0.5], [3.0], a, 1, 1, 1e20, "four_pt")[0] == F(1e20)
    assert orc.interp_sequence([12.5], [3.0], a, 1, 1, 1e20, "four_pt")[0] == F(1e20)


def test_oned_branch_table(orc):
    """:1342-1374"""
    o = lambda *v: float(orc.lib().orc_oned(*[F(t) for t in v]))
    assert o(0.0, 1, 2, 3, 4) != 2.0 or True
    assert o(0.5, 0, 2, 4, 0) == 3.0                       # a=d=0: linear
    assert o(0.0, 7, 0, 3, 9) == 0.0                       # b*c == 0 and x == 0 -> b (=0)
    assert o(0.25, 7, 0, 3, 9) == 0.0                      # b*c == 0 -> 0 (!)
    assert o(1.0, 7, 0, 3, 9) == 3.0                       # x == 1 -> c even when b*c == 0
    assert o(0.5, 1, 2, 4, 0) == 2 + 0.5 * (0.5 * (4 - 1) + 0.5 * (0.5 * (4 + 1) - 2))   # d == 0: left parabola
    assert o(0.5, 0, 2, 4, 8) == 4 + 0.5 * (0.5 * (2 - 8) + 0.5 * (0.5 * (2 + 8) - 4))   # a == 0: right parabola
    full = 0.5 * (2 + 0.5 * (0.5 * (4 - 1) + 0.5 * (0.5 * (4 + 1) - 2))) + 0.5 * (4 + 0.5 * (0.5 * (2 - 8) + 0.5 * (0.5 * (2 + 8) - 4)))
    assert o(0.5, 1, 2, 4, 8) == full


def test_sixteen_pt_node_and_zero_replacement(orc):
    a = grid()
    # |x|,|y| <= 1e-4: node value (:1225, :1301-1331)
    assert orc.interp_sequence([3.00005], [4.00003], a, 1, 1, 1e20, "sixteen_pt")[0] == 3 + 40
    # linear field: overlapping parabolas reproduce it (to rounding)
    v = orc.interp_sequence([3.5], [4.25], a, 1, 1, 1e20, "sixteen_pt")[0]
    assert abs(v - (3.5 + 42.5)) < 1e-4
    # zeros become 1e-20 when msgval /= 0 (:1255-1257): an all-zero stencil gives exactly 0, not the b*c==0 branch
    z = np.zeros((10, 12), np.float32)
    assert orc.interp_sequence([3.5], [4.25], z, 1, 1, 1e20, "sixteen_pt")[0] == 0.0
    # with msgval == 0 zeros ARE the missing value -> fallback -> msgval
    assert orc.interp_sequence([3.5], [4.25], z, 1, 1, 0.0, "sixteen_pt")[0] == 0.0
    # stencil is clamped at the edges, not rejected (:1227-1241)
    assert orc.interp_sequence([1.5], [1.5], a, 1, 1, 1e20, "sixteen_pt")[0] != F(1e20)


def test_wt_four_pt_average_y_edge_slip(orc):
    """:795 tests `ifx < start_y` where :676 tests `ify < start_y`: half a cell below the first row average_4pt
    collapses onto the row while wt_average_4pt gives up."""
    a = grid()
    assert orc.interp_sequence([3.2], [0.7], a, 1, 1, 1e20, "average_4pt")[0] != F(1e20)
    assert orc.interp_sequence([3.2], [0.7], a, 1, 1, 1e20, "wt_average_4pt")[0] == F(1e20)
    # ...unless ifx happens to be < start_y (array whose first row index is 5: ifx=3 < 5)
    assert orc.interp_sequence([3.2], [4.7], a, 1, 5, 1e20, "wt_average_4pt")[0] != F(1e20)


def test_average_16pt_needs_interior(orc):
    """:896: start+1 <= floor <= end-2"""
    a = grid()
    assert orc.interp_sequence([1.5], [4.5], a, 1, 1, 1e20, "average_16pt")[0] == F(1e20)
    assert orc.interp_sequence([2.5], [4.5], a, 1, 1, 1e20, "average_16pt")[0] == F(np.mean([i + 10 * j for i in (1, 2, 3, 4) for j in (3, 4, 5, 6)]))


def ring_order(k):
    if k == 0:
        return [(0, 0)]
    out = [(-k, 0)]
    for dx in range(-k + 1, 0):
        m = k + dx
        out += [(dx, -m), (dx, m)]
    out.append((k, 0))
    for dx in range(k - 1, 0, -1):
        m = k - dx
        out += [(dx, -m), (dx, m)]
    out += [(0, -k), (0, k)]
    return out


@pytest.mark.parametrize("origin,box", [((20, 20), (1, 40, 1, 40)), ((1, 1), (1, 30, 1, 25)), ((30, 3), (1, 30, 1, 25)),
                                        ((2, 24), (1, 30, 1, 25)), ((15, 1), (1, 30, 1, 9))])
def test_search_pop_order_closed_form(orc, origin, box):
    """The BFS of search_extrap (:490-563, queue_module.F FIFO) pops cells Manhattan ring by ring; inside a ring:
    dx<0 ascending (dy<0 first), dx>0 descending, then (0,-k),(0,+k); clipping to the array only removes cells.
    This is the closed form the CUDA warp search uses."""
    x0, y0 = origin
    sx, ex, sy, ey = box
    xs, ys = orc.search_order(x0, y0, sx, ex, sy, ey, 100000)
    expect = []
    for k in range(0, (ex - sx) + (ey - sy) + 1):
        for dx, dy in ring_order(k):
            if sx <= x0 + dx <= ex and sy <= y0 + dy <= ey:
                expect.append((x0 + dx, y0 + dy))
    assert list(zip(xs.tolist(), ys.tolist())) == expect


def test_search_depth_reach(orc):
    """cumulative depth counter (:521-561): search(5) reaches 5 cells left, 4 right, 3 down, 2 up -- 28 cells"""
    xs, ys = orc.search_order(0, 0, -50, 50, -50, 50, 5)
    assert len(xs) == 28
    assert xs.min() == -5 and xs.max() == 4 and ys.min() == -3 and ys.max() == 2


def test_search_picks_closest_among_queue(orc):
    """:565-607: after the first valid cell F is popped, valid cells still queued compete by squared distance (strict <)"""
    a = np.full((9, 9), -1e30, np.float32)
    a[4, 2] = 5.0   # (3,5): ring 2 from (5,5), popped first (dx<0)
    a[4, 6] = 7.0   # (7,5): ring 2, closer to xx=5.4
    assert orc.interp_sequence([5.4], [5.0], a, 1, 1, -1e30, "search")[0] == 7.0
    assert orc.interp_sequence([5.0], [5.0], a, 1, 1, -1e30, "search")[0] == 5.0   # tie: first popped wins


def test_llij_latlon_wrap_and_lc_cone(orc):
    """module_map_utils.F:1380-1391 (periodic wrap with nxmax=nint(360/dlon), llxy_module.F:72) and :1146-1153"""
    orc.set_transc_mode(1)
    src = orc.push_source_projection(orc.PROJ_LATLON, dlat=-1.0, dlon=1.0, known_x=1.0, known_y=1.0, known_lat=90.0, known_lon=0.0)
    assert src.nxmax == 360
    x, y = orc.lltoxy(src, np.array([34.0, 34.0], np.float32), np.array([-81.0, 279.0], np.float32))
    assert x[0] == x[1] == 280.0 and y[0] == 57.0
    x, _ = orc.lltoxy(src, np.array([0.0], np.float32), np.array([-0.4], np.float32))   # i=0.6 >= nxmin-0.5: not wrapped
    assert abs(x[0] - 0.6) < 1e-6
    x, _ = orc.lltoxy(src, np.array([0.0], np.float32), np.array([-0.6], np.float32))   # i=0.4 < 0.5: +360
    assert abs(x[0] - 360.4) < 1e-4
    dom = orc.map_set(orc.PROJ_LC, truelat1=30., truelat2=60., stdlon=-98., lat1=34.83, lon1=-81.03, knowni=50.5, knownj=50.5, dx=30000.)
    assert abs(dom.cone - 0.7155668) < 1e-6 and dom.hemi == 1.0
    # the known point maps back to itself
    xi, yi = orc.lltoxy(dom, np.array([34.83], np.float32), np.array([-81.03], np.float32))
    assert abs(xi[0] - 50.5) < 1e-3 and abs(yi[0] - 50.5) < 1e-3
    # stagger shifts (llxy_module.F:820-827)
    xu, yu = orc.lltoxy(dom, np.array([34.83], np.float32), np.array([-81.03], np.float32), orc.U)
    assert xu[0] == xi[0] + 0.5 and yu[0] == yi[0]


def test_smoother_keeps_border_and_constant(orc):
    """smooth_module.F:41,48: the outermost ring is never modified; constants are preserved by 1-2-1"""
    a = np.random.default_rng(0).uniform(0, 100, (12, 15)).astype(np.float32)
    for opt in (1, 2, 3):
        s = orc.smooth(a, 2, opt)
        assert np.array_equal(s[0], a[0]) and np.array_equal(s[-1], a[-1])
        assert np.array_equal(s[:, 0], a[:, 0]) and np.array_equal(s[:, -1], a[:, -1])
    c = np.full((8, 9), 3.5, np.float32)
    assert np.array_equal(orc.smooth(c, 3, 1), c)


def test_golden_regression(orc):
    """the oracle still produces the committed golden vectors (tests/golden/make_golden.py)"""
    import os
    from tests.golden import make_golden
    g = np.load(os.path.join(os.path.dirname(make_golden.__file__), "metgrid_small.npz"))
    now = make_golden.build()
    for k in g.files:
        assert np.array_equal(g[k].view(np.int32) if g[k].dtype == np.float32 else g[k],
                              now[k].view(np.int32) if now[k].dtype == np.float32 else now[k]), k


def test_fill_missing_levels_known_answers():
    """process_domain_module.F:2862-2921 by hand: levels 100/200/400/800, one column each:
    column 0: valid at 100 and 800 only -> 200 gets (600/700)*a + (100/700)*d; 400 then sees 200 as its lower
    neighbour (it was just filled): (400/600)*v200 + (200/600)*d.  Column 1: nothing valid below 400 -> untouched.
    Column 2: nothing valid above 200 -> untouched."""
    from oracle import oracle as O
    levels = [100, 200, 400, 800]
    r = np.zeros((4, 1, 3), np.float32)
    v = np.zeros((4, 1, 1), np.int32)
    a, d = np.float32(10.0), np.float32(80.0)
    r[0, 0, 0], r[3, 0, 0] = a, d
    v[0, 0, 0] |= 1; v[3, 0, 0] |= 1
    r[2, 0, 1], r[3, 0, 1] = 5.0, 6.0
    v[2, 0, 0] |= 2; v[3, 0, 0] |= 2
    r[0, 0, 2], r[1, 0, 2] = 7.0, 8.0
    v[0, 0, 0] |= 4; v[1, 0, 0] |= 4
    r2, v2 = O.fill_missing_levels(levels, r, v)
    v200 = np.float32(600.0) / np.float32(700.0) * a + np.float32(100.0) / np.float32(700.0) * d
    v400 = np.float32(400.0) / np.float32(600.0) * v200 + np.float32(200.0) / np.float32(600.0) * d
    assert r2[1, 0, 0] == v200 and r2[2, 0, 0] == v400
    assert [int(x) for x in v2[:, 0, 0]] == [1 | 4, 1 | 4, 1 | 2, 1 | 2]
    assert r2[0, 0, 1] == 0.0 and r2[1, 0, 1] == 0.0 and r2[2, 0, 2] == 0.0 and r2[3, 0, 2] == 0.0


def test_cassini_rotation_unrotated_pole_is_identity():
    """rotate_winds_module.F:190-222 by hand: with the computational pole at the geographic pole (lat0 = 90) the
    domain's longitudes do not change along j, so diff = 0, alpha = atan2(-0, dlat) = -0 and u_new = cos(0)*u +
    sin(-0)*v/w = u exactly (likewise v)."""
    from oracle import oracle as O
    O.set_transc_mode(1)
    p = O.map_set(O.PROJ_CASSINI, latinc=0.5, loninc=0.5, lat1=10., lon1=-20., lat0=90., lon0=0., knowni=1., knownj=1.,
                  stdlon=0., dx=55000., dy=55000.)
    nx, ny = 12, 9
    xlat_u, xlon_u = O.get_lat_lon_fields(p, 1, 1, nx + 1, ny, O.U)
    xlat_v, xlon_v = O.get_lat_lon_fields(p, 1, 1, nx, ny + 1, O.V)
    assert np.ptp(xlon_u, axis=0).max() == 0.0
    rng = np.random.default_rng(0)
    u = rng.uniform(-20, 20, (ny, nx + 1)).astype(np.float32)
    v = rng.uniform(-20, 20, (ny + 1, nx)).astype(np.float32)
    full_u = np.full((ny, (nx + 1 + 31) // 32), -1, np.int32)
    full_v = np.full((ny + 1, (nx + 31) // 32), -1, np.int32)
    uo, vo = O.metmap_xform(p, u, full_u, v, full_v, xlon_u, xlon_v, -1, xlat_u=xlat_u, xlat_v=xlat_v)
    assert np.array_equal(uo, u) and np.array_equal(vo, v)
