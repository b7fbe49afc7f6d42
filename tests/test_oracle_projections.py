"""CPU: pins for the map-projection restatement (oracle/orc_proj.c <- module_map_utils.F) that do not depend on it:
(1) textbook float64 formulas of the Lambert conformal, polar stereographic and Mercator projections (Snyder 1987,
eqs. 15-1..15-4, 21-5..21-8, 7-1..7-2 on the sphere) evaluated with numpy, against orc_lltoxy; (2) the round trip
xytoll(lltoxy(p)) for every supported projection; (3) the known point maps to (knowni, knownj)."""
import numpy as np
import pytest

R = 6370000.0      # EARTH_RADIUS_M, constants_module.F


def lc_xy(lat, lon, truelat1, truelat2, stdlon, lat1, lon1, knowni, knownj, dx):
    rad = np.pi / 180.0
    hemi = 1.0 if truelat1 >= 0 else -1.0
    p1, p2 = hemi * truelat1 * rad, hemi * truelat2 * rad
    if abs(truelat1 - truelat2) > 0.1:
        n = np.log(np.cos(p1) / np.cos(p2)) / np.log(np.tan(np.pi / 4 + p2 / 2) / np.tan(np.pi / 4 + p1 / 2))
    else:
        n = np.sin(p1)
    F = np.cos(p1) * np.tan(np.pi / 4 + p1 / 2) ** n / n

    def xy(la, lo):
        rho = R * F / np.tan(np.pi / 4 + hemi * la * rad / 2) ** n
        d = lo - stdlon
        d = (d + 180.0) % 360.0 - 180.0
        th = n * d * rad
        return hemi * rho * np.sin(th), -rho * np.cos(th)     # y measured from the pole, positive towards it
    x0, y0 = xy(lat1, lon1)
    x, y = xy(lat, lon)
    return knowni + hemi * (x - x0) / dx, knownj + hemi * (y - y0) / dx


def ps_xy(lat, lon, truelat1, stdlon, lat1, lon1, knowni, knownj, dx):
    rad = np.pi / 180.0
    hemi = 1.0 if truelat1 >= 0 else -1.0

    def xy(la, lo):
        rho = R * (1.0 + np.sin(hemi * truelat1 * rad)) * np.tan(np.pi / 4 - hemi * la * rad / 2)
        d = (lo - stdlon) * rad
        return rho * np.sin(d), -hemi * rho * np.cos(d)
    x0, y0 = xy(lat1, lon1)
    x, y = xy(lat, lon)
    return knowni + (x - x0) / dx, knownj + (y - y0) / dx


def merc_xy(lat, lon, truelat1, lat1, lon1, knowni, knownj, dx):
    rad = np.pi / 180.0
    k = R * np.cos(truelat1 * rad)
    d = (lon - lon1 + 180.0) % 360.0 - 180.0
    x = k * d * rad
    y = k * (np.log(np.tan(np.pi / 4 + lat * rad / 2)) - np.log(np.tan(np.pi / 4 + lat1 * rad / 2)))
    return knowni + x / dx, knownj + y / dx


CASES = [
    ("lc", lc_xy, dict(truelat1=30., truelat2=60., stdlon=-98., lat1=34.83, lon1=-81.03, knowni=50.5, knownj=50.5, dx=30000.), (20, 55, -125, -60)),
    ("lc_tangent", lc_xy, dict(truelat1=38.5, truelat2=38.5, stdlon=-97.5, lat1=38.5, lon1=-97.5, knowni=900., knownj=530., dx=3000.), (22, 52, -125, -65)),
    ("lc_south", lc_xy, dict(truelat1=-30., truelat2=-60., stdlon=140., lat1=-35., lon1=145., knowni=40., knownj=40., dx=12000.), (-50, -20, 120, 165)),
    ("ps", ps_xy, dict(truelat1=60., stdlon=-100., lat1=70., lon1=-110., knowni=60., knownj=60., dx=25000.), (45, 88, -180, 180)),
    ("ps_south", ps_xy, dict(truelat1=-71., stdlon=0., lat1=-75., lon1=20., knowni=50., knownj=50., dx=30000.), (-88, -50, -180, 180)),
    ("merc", merc_xy, dict(truelat1=10., lat1=5., lon1=100., knowni=30., knownj=30., dx=27000.), (-30, 40, 70, 150)),
]


@pytest.mark.parametrize("name,fn,kw,box", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("mode", [0, 1])
def test_forward_transform_matches_textbook_formulas(orc, name, fn, kw, box, mode):
    orc.set_transc_mode(mode)
    code = {"lc": orc.PROJ_LC, "ps": orc.PROJ_PS, "merc": orc.PROJ_MERC}[name.split("_")[0]]
    p = orc.map_set(code, **kw)
    rng = np.random.default_rng(1)
    lat = rng.uniform(box[0], box[1], 4000).astype(np.float32)
    lon = rng.uniform(box[2], box[3], 4000).astype(np.float32)
    x, y = orc.lltoxy(p, lat, lon)
    xr, yr = fn(lat.astype(np.float64), lon.astype(np.float64), **kw)
    # binary32 evaluation: the pole is hundreds to thousands of cells away, so a few ulp of that distance remain --
    # 5e-4 cell per 100 cells of index magnitude (15 m on the 30-km grid), far below anything a formula error would give
    tol = 5e-4 * np.maximum(1.0, np.maximum(np.abs(xr), np.abs(yr)) / 100.0)
    assert np.all(np.abs(x - xr) < tol), (name, np.max(np.abs(x - xr)))
    assert np.all(np.abs(y - yr) < tol), (name, np.max(np.abs(y - yr)))
    # the known point sits on (knowni, knownj)
    xk, yk = orc.lltoxy(p, np.float32([kw["lat1"]]), np.float32([kw["lon1"]]))
    assert abs(xk[0] - kw["knowni"]) < 2e-3 and abs(yk[0] - kw["knownj"]) < 2e-3


ALL = {
    "lc": ("PROJ_LC", dict(truelat1=30., truelat2=60., stdlon=-98., lat1=34.83, lon1=-81.03, knowni=50.5, knownj=50.5, dx=30000.), (20, 55, -125, -60)),
    "ps": ("PROJ_PS", dict(truelat1=60., stdlon=-100., lat1=70., lon1=-110., knowni=60., knownj=60., dx=25000.), (45, 88, -179, 179)),
    "merc": ("PROJ_MERC", dict(truelat1=10., lat1=5., lon1=100., knowni=30., knownj=30., dx=27000.), (-30, 40, 70, 150)),
    "latlon": ("PROJ_LATLON", dict(latinc=-0.25, loninc=0.25, knowni=1., knownj=1., lat1=90., lon1=0., nxmax=1440), (-89, 89, 0, 359)),
    "cassini": ("PROJ_CASSINI", dict(latinc=0.5, loninc=0.5, lat1=10., lon1=-20., lat0=60., lon0=-140., knowni=1., knownj=1., stdlon=-140., dx=55000., dy=55000.), (15, 60, -10, 60)),
    "cyl": ("PROJ_CYL", dict(latinc=1.0, loninc=1.0, stdlon=0.), (-80, 80, 1, 359)),
    "ps_wgs84": ("PROJ_PS_WGS84", dict(truelat1=70., stdlon=-45., lat1=65., lon1=-50., knowni=100., knownj=100., dx=5000.), (60, 85, -80, -10)),
    "albers": ("PROJ_ALBERS_NAD83", dict(truelat1=29.5, truelat2=45.5, stdlon=-96., lat1=23., lon1=-96., knowni=1., knownj=1., dx=1000.), (25, 48, -120, -70)),
}


@pytest.mark.parametrize("name", list(ALL))
def test_round_trip_every_projection(orc, name):
    """xytoll(lltoxy(lat, lon)) returns the point (to binary32 accuracy of the intermediate index)."""
    code, kw, box = ALL[name]
    orc.set_transc_mode(1)
    p = orc.map_set(getattr(orc, code), **kw)
    rng = np.random.default_rng(2)
    lat = rng.uniform(box[0], box[1], 3000).astype(np.float32)
    lon = rng.uniform(box[2], box[3], 3000).astype(np.float32)
    x, y = orc.lltoxy(p, lat, lon)
    lat2, lon2 = orc.xytoll(p, x, y)
    dlon = (lon2 - lon + 180.0) % 360.0 - 180.0
    coslat = np.maximum(np.cos(np.deg2rad(lat)), 0.02)
    tol = 2e-3 if name in ("albers", "ps_wgs84", "ps") else 5e-4      # degrees; fine grids and points next to the pole amplify the binary32 index error
    assert np.max(np.abs(lat2 - lat)) < tol, (name, np.max(np.abs(lat2 - lat)))
    assert np.max(np.abs(dlon) * coslat) < tol, (name, np.max(np.abs(dlon) * coslat))
