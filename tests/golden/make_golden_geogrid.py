"""Regenerates tests/golden/geogrid_small.npz from the CPU oracle (FP64-then-round transcendental mode): a 21-category
land-use accumulation (four 60x60 tiles onto a 9-km Lambert patch) with fractions / dominant category / landmask, a
continuous accumulation with the point fallback, the three smoothers and a Lambert wind rotation.  Like
metgrid_small.npz these fixtures pin the ORACLE (the reference ships no golden vectors and cannot be run here); the GPU
tests hold the product to them.  Run from the repo root:  python tests/golden/make_golden_geogrid.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

NX, NY, DX, DXDEG, TN = 36, 30, 9000.0, 0.00833333, 60
KW_DOM = dict(truelat1=30.0, truelat2=60.0, stdlon=-98.0, lat1=39.0, lon1=-105.0, knowni=(NX + 1) / 2.0,
              knownj=(NY + 1) / 2.0, dx=DX)
KW_SRC = dict(dlat=DXDEG, dlon=DXDEG, known_x=1.0, known_y=1.0, known_lat=-89.99583, known_lon=-179.99583)


def cat_tile(tx, ty, tn, ncat=21):
    j, i = np.mgrid[ty:ty + tn, tx:tx + tn]
    patch = ((i // 7) * 31 + (j // 5) * 17) % ncat + 1
    noise = (i * 7919 + j * 104729) % 13 == 0
    return np.where(noise, (i + j) % ncat + 1, patch).astype(np.float32)[None]


def topo_tile(tx, ty, tn):
    j, i = np.mgrid[ty:ty + tn, tx:tx + tn]
    return np.round(1500 + 1200 * np.sin(i / 37.0) * np.cos(j / 53.0) + ((i * 13 + j * 7) % 17)).astype(np.float32)[None]


def tile_origins(xlat, xlon, bdr):
    ix0 = int(np.floor((float(xlon.min()) - 0.05 - KW_SRC["known_lon"]) / DXDEG)) + 1
    ix1 = int(np.ceil((float(xlon.max()) + 0.05 - KW_SRC["known_lon"]) / DXDEG)) + 1
    iy0 = int(np.floor((float(xlat.min()) - 0.05 - KW_SRC["known_lat"]) / DXDEG)) + 1
    iy1 = int(np.ceil((float(xlat.max()) + 0.05 - KW_SRC["known_lat"]) / DXDEG)) + 1
    out = []
    for ty in range(TN * ((iy0 - 1) // TN) + 1, iy1 + 1, TN):
        for tx in range(TN * ((ix0 - 1) // TN) + 1, ix1 + 1, TN):
            out.append((tx - bdr, ty - bdr))
    return out


def accumulate(itype_cat, xlat, xlon, dom, src, chain, nk, bdr, tile_fn):
    """calc_field for one priority level with the oracle: accumulate + per-point fallback per tile, then finish."""
    gy, gx = xlat.shape
    field = np.zeros((nk, gy, gx), np.float32)
    count = np.zeros_like(field)
    nxw = (gx + 31) // 32
    processed = np.zeros((gy, nxw), np.int32)
    level = np.zeros((gy, nxw), np.int32)
    itype = 1 if itype_cat else 2   # wpsgpu.h: CATEGORICAL = 1, SP_CONTINUOUS = 2 (same codes as the oracle)
    for tx, ty in tile_origins(xlat, xlon, bdr):
        tile = tile_fn(tx, ty, TN + 2 * bdr)
        if itype_cat:
            O.accum_categorical(src, dom, O.M, tile[0], tx, ty, bdr, field, -2, -2, 1, processed, level, 1, 1e20)
        else:
            O.accum_continuous(src, dom, O.M, tile, tx, ty, 1, bdr, field, count, -2, -2, 1, processed, level, 1e20)
        O.tile_point_fallback(src, itype, tile, tx, ty, 1, bdr, xlat, xlon, None, False, -1.0, chain, 1e20, 1e20, field,
                              count, -2, -2, 1, processed, level)
    O.calc_field_finish(itype, field, count, None, False, -1.0, 1e20, None, -2, -2, 1, processed, level)
    return field


def build():
    O.set_transc_mode(1)
    dom = O.map_set(O.PROJ_LC, **KW_DOM)
    src = O.push_source_projection(O.PROJ_LATLON, **KW_SRC)
    gx, gy = NX + 6, NY + 6
    xlat, xlon = O.get_lat_lon_fields(dom, -2, -2, gx, gy, O.M)
    out = dict(xlat=xlat, xlon=xlon)
    counts = accumulate(True, xlat, xlon, dom, src, "nearest_neighbor", 21, 0, lambda tx, ty, tn: cat_tile(tx, ty, tn))
    out["landuse_counts"] = counts
    fr = O.fractions(counts, 1e20)
    out["landuse_fractions"] = fr
    out["landuse_dominant_plain"] = O.dominant_category(fr, 1, 1e20)                      # process_tile_module.F:1123-1144
    out["landmask"] = O.landmask_from_fractions(fr, 1, [17, 21], True, 1e20)              # :570-596
    out["landuse_dominant"] = O.dominant_category_landmask(fr, out["landmask"], 1, [17, 21], True)   # :699-727 (LU_INDEX)
    hgt = accumulate(False, xlat, xlon, dom, src, "average_gcell(4.0)+four_pt+average_4pt", 1, 3, lambda tx, ty, tn: topo_tile(tx, ty, tn))
    out["hgt"] = hgt
    for name, opt, npass in (("hgt_121_2", 1, 2), ("hgt_smth_desmth_1", 2, 1), ("hgt_smth_desmth_special_2", 3, 2)):
        out[name] = O.smooth(hgt[0] - 1600.0, npass, opt)
    # Lambert wind rotation on the same domain's U / V staggers, every mask bit set
    _, xlon_u = O.get_lat_lon_fields(dom, 1, 1, NX + 1, NY, O.U)
    _, xlon_v = O.get_lat_lon_fields(dom, 1, 1, NX, NY + 1, O.V)
    rng = np.random.default_rng(11)
    u = rng.uniform(-25, 25, (NY, NX + 1)).astype(np.float32)
    v = rng.uniform(-25, 25, (NY + 1, NX)).astype(np.float32)
    um = np.full((NY, (NX + 1 + 31) // 32), -1, np.int32)
    vm = np.full((NY + 1, (NX + 31) // 32), -1, np.int32)
    ur, vr = O.metmap_xform(dom, u, um, v, vm, xlon_u, xlon_v, 1)
    out.update(u=u, v=v, xlon_u=xlon_u, xlon_v=xlon_v, u_rot=ur, v_rot=vr)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "geogrid_small.npz"), **build())
    print("wrote geogrid_small.npz")
