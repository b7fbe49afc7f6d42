"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from oracle import oracle as O
from wps_b200 import capi


def conus_like(nx, ny, dx, truelat1=38.5, truelat2=38.5, stand_lon=-97.5, ref_lat=38.5, ref_lon=-97.5):
    """(oracle proj, product proj) of a Lambert domain centred on (ref_lat, ref_lon)."""
    kw = dict(truelat1=truelat1, truelat2=truelat2, stdlon=stand_lon, lat1=ref_lat, lon1=ref_lon,
              knowni=(nx + 1) / 2.0, knownj=(ny + 1) / 2.0, dx=dx)
    return O.map_set(O.PROJ_LC, **kw), capi.map_set(capi.PROJ_LC, **kw)


def latlon_source(nx, ny, dlat, dlon, startlat, startlon, earth_radius=6371229.0):
    kw = dict(dlat=dlat, dlon=dlon, known_x=1.0, known_y=1.0, known_lat=startlat, known_lon=startlon,
              earth_radius=earth_radius)
    return O.push_source_projection(O.PROJ_LATLON, **kw), capi.push_source_projection(capi.PROJ_LATLON, **kw)


def halo(slab, bdr=3):
    """halo_slab of process_domain_module.F:1121-1135 for a global lat-lon source."""
    ny, nx = slab.shape
    out = np.empty((ny, nx + 2 * bdr), np.float32)
    out[:, bdr:bdr + nx] = slab
    out[:, :bdr] = slab[:, nx - bdr:]
    out[:, bdr + nx:] = slab[:, :bdr]
    return out


def synth_field(nx, ny, seed, kind="smooth", missing=None, frac_missing=0.0, zeros=0.0):
    rng = np.random.default_rng(seed)
    lam = np.linspace(0, 2 * np.pi, nx, endpoint=False, dtype=np.float64)[None, :]
    phi = np.linspace(np.pi / 2, -np.pi / 2, ny, dtype=np.float64)[:, None]
    if kind == "smooth":
        f = 280.0 + 30.0 * np.sin(3 * lam) * np.cos(2 * phi) + rng.uniform(-0.03, 0.03, (ny, nx))
    elif kind == "wind":
        f = 12.0 * np.sin(2 * lam + 0.3) * np.cos(3 * phi) + rng.uniform(-0.01, 0.01, (ny, nx))
    else:
        f = rng.uniform(-5, 5, (ny, nx))
    f = f.astype(np.float32)
    if zeros > 0:
        f[rng.random((ny, nx)) < zeros] = 0.0
    if missing is not None and frac_missing > 0:
        f[rng.random((ny, nx)) < frac_missing] = missing
    return f


def landsea(nx, ny, seed=0, thresh=0.3):
    lam = np.linspace(0, 2 * np.pi, nx, endpoint=False)[None, :]
    phi = np.linspace(np.pi / 2, -np.pi / 2, ny)[:, None]
    return (np.sin(5 * lam + seed) + np.cos(4 * phi) > thresh).astype(np.float32)


def bits_to_bool(words, nx):
    ny = words.shape[0]
    b = ((words.astype(np.uint32)[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).astype(bool)
    return b.reshape(ny, -1)[:, :nx]


def bool_to_bits(mask):
    ny, nx = mask.shape
    nw = (nx + 31) // 32
    pad = np.zeros((ny, nw * 32), np.uint32)
    pad[:, :nx] = mask
    w = (pad.reshape(ny, nw, 32) << np.arange(32, dtype=np.uint32)[None, None, :]).sum(axis=2, dtype=np.uint64)
    return w.astype(np.uint32).view(np.int32)


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7fffffff), a)
    b = np.where(b < 0, -(b & 0x7fffffff), b)
    return np.abs(a - b)


def assert_bitexact(a, b, what=""):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = a.view(np.int32) == b.view(np.int32)
    if not same.all():
        bad = np.argwhere(~same)
        k = tuple(bad[0])
        raise AssertionError("%s: %d of %d values differ; first at %s: %r vs %r" %
                             (what, (~same).sum(), same.size, k, a[k], b[k]))


def assert_sixteen_pt_tolerance(r, r_ref, scale, what="", rtol=1e-5):
    """Tolerance mode of the device path (wpsgpu_metgrid_set_precision): north_star allows 1e-5 relative for
    sixteen_pt.  `scale` is the largest magnitude among the usable source values: the reference's own evaluation of
    the overlapping parabolas carries a rounding error of a few ulp of the stencil values, so a result that is small
    only through cancellation cannot be pinned more tightly than rtol * scale.  Magnitudes below 1e-19 are the
    reference's stand-ins for zero (interp_module.F:1255-1257, :1317) and compare equal to 0.
    Returns (share of points within a pure relative 1e-5, largest error / scale)."""
    a = np.ascontiguousarray(r).astype(np.float64)
    b = np.ascontiguousarray(r_ref).astype(np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = rtol * np.maximum(np.abs(b), scale) + 1e-19
    bad = err > tol
    if bad.any():
        k = tuple(np.argwhere(bad)[0])
        raise AssertionError("%s: %d of %d values outside %g * max(|ref|, %g); first at %s: %r vs %r" %
                             (what, bad.sum(), bad.size, rtol, scale, k, r[k], r_ref[k]))
    pure = err <= rtol * np.abs(b) + 1e-19
    return float(pure.mean()), float((err / max(scale, 1e-30)).max())
