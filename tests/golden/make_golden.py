"""Regenerates tests/golden/metgrid_small.npz from the CPU oracle (FP64-then-round transcendental mode).
The reference ships no golden vectors, and cannot be run here, so these fixtures pin the
ORACLE's behaviour (and through the GPU tests, the product's) against accidental change; they are not outputs of
the Fortran itself.  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.util import halo, landsea, synth_field  # noqa: E402

CASES = [
    ("c3d", "sixteen_pt+four_pt+average_4pt", 1e20, 0.0, None),
    ("psfc", "four_pt+average_4pt", 1e20, 1e20, None),
    ("landsea", "nearest_neighbor", 1e20, -1.0, None),
    ("soil", "sixteen_pt+four_pt+wt_average_4pt+wt_average_16pt+search", -1e30, 285.0, 0.0),
    ("sst", "sixteen_pt+four_pt", -1e30, 0.0, None),
    ("avg16", "average_16pt+wt_average_16pt+search(6)", -1e30, 1e20, 1.0),
]


def build():
    O.set_transc_mode(1)
    nx, ny = 40, 30
    dom = O.map_set(O.PROJ_LC, truelat1=30.0, truelat2=60.0, stdlon=-98.0, lat1=34.83, lon1=-81.03,
                    knowni=(nx + 1) / 2.0, knownj=(ny + 1) / 2.0, dx=60000.0)
    xlat, xlon = O.get_lat_lon_fields(dom, 1, 1, nx, ny, O.M)
    src = O.push_source_projection(O.PROJ_LATLON, dlat=-1.0, dlon=1.0, known_x=1.0, known_y=1.0, known_lat=90.0,
                                   known_lon=0.0, earth_radius=6371229.0)
    ls = halo(landsea(360, 181))
    out = dict(xlat=xlat, xlon=xlon, landsea=ls)
    for k, (name, chain, msg, fill, mval) in enumerate(CASES):
        slab = halo(synth_field(360, 181, 1234 + k, missing=np.float32(msg), frac_missing=0.15 if msg < 0 else 0.0, zeros=0.01))
        mv = (mval, 1e20, 1e20) if mval is not None else (1e20, 1e20, 1e20)
        fo = O.field_opts(chain, np.float32(msg), np.float32(fill), -2.0, mv)
        r, npw = O.interp_met_field(src, fo, O.M, xlat, xlon, 1, 1, (1, nx, 1, ny), slab, -2, 1, 3,
                                    mask=ls if mval is not None else None)
        out[name + "_slab"] = slab
        out[name + "_r"] = r
        out[name + "_new_pts"] = npw
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "metgrid_small.npz"), **build())
    print("wrote metgrid_small.npz")
