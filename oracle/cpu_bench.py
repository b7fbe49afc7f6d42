"""CPU baseline driver (test infrastructure): times the oracle's interp_met_field on a bounded sample of the
bench workload, patch-decomposed over host processes the way WPS's dmpar build does (parallel_module.F:62-79:
near-square nproc_x x nproc_y, every process reads the whole source slab, no output gather timed).
The real metgrid.exe cannot be built in this image (no Fortran compiler / MPI / netCDF / WRF I/O), so this
restatement is the reference arm; it recomputes lltoxy for every slab exactly like the reference does."""
import multiprocessing as mp
import os
import time

import numpy as np

from . import oracle as O


def proc_grid(nprocs):
    """parallel_module.F:62-79: nproc_x*nproc_y = nprocs, as square as possible."""
    best = (1, nprocs)
    for px in range(1, nprocs + 1):
        if nprocs % px == 0:
            py = nprocs // px
            if abs(px - py) < abs(best[0] - best[1]):
                best = (px, py)
    return best


def patches(nx, ny, nprocs):
    px, py = proc_grid(nprocs)
    xs = np.linspace(0, nx, px + 1).astype(int)
    ys = np.linspace(0, ny, py + 1).astype(int)
    return [(xs[i] + 1, xs[i + 1], ys[j] + 1, ys[j + 1]) for j in range(py) for i in range(px)]


_G = {}


def _worker(args):
    patch, idxs = args
    src = O.push_source_projection(O.PROJ_LATLON, **_G["src_kw"])
    t0 = time.perf_counter()
    for k in idxs:
        s = _G["sample"][k]
        xlat, xlon = _G["grids"][s["stagger"]]
        fo = O.field_opts(s["chain"], np.float32(s["missing"]), np.float32(s["fill"]), s["masked"], s["mask_val"])
        O.interp_met_field(src, fo, s["stagger"], xlat, xlon, 1, 1, _G["sd"], s["slab"], -2, 1, 3, mask=s["mask"],
                           land_mask=s["land_mask"], water_mask=s["water_mask"],
                           landmask=_G["landmask"] if s["use_landmask"] else None, sub=patch)
    return time.perf_counter() - t0


def run(src_kw, grids, landmask, sd, sample, nprocs=None, repeats=1):
    """-> (points_per_second, seconds, points, nprocs).  grids: {stagger: (xlat, xlon)} numpy arrays."""
    O.lib()
    O.set_transc_mode(0)  # glibc float libm, as gfortran would call
    nprocs = nprocs or os.cpu_count() or 1
    _G.update(src_kw=src_kw, grids=grids, landmask=landmask, sd=sd, sample=sample)
    pts = sum(grids[s["stagger"]][0].size for s in sample) * repeats
    nx, ny = sd[1] - sd[0] + 1, sd[3] - sd[2] + 1
    # each process owns one patch of the mass grid (staggered grids get the extra row/column on the last patch)
    pat = patches(nx, ny, nprocs)
    jobs = []
    px, py = proc_grid(nprocs)
    for n, (a, b, c, d) in enumerate(pat):
        jobs.append(((a, b + (1 if (n % px) == px - 1 else 0), c, d + (1 if (n // px) == py - 1 else 0)), list(range(len(sample))) * repeats))
    # clip per slab inside the worker: oracle clips by array bounds
    t0 = time.perf_counter()
    if nprocs == 1:
        _worker_clip(jobs[0])
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(nprocs) as pool:
            pool.map(_worker_clip, jobs)
    dt = time.perf_counter() - t0
    return pts / dt, dt, pts, nprocs


def _worker_clip(args):
    (a, b, c, d), idxs = args
    src = O.push_source_projection(O.PROJ_LATLON, **_G["src_kw"])
    for k in idxs:
        s = _G["sample"][k]
        xlat, xlon = _G["grids"][s["stagger"]]
        ny, nx = xlat.shape
        patch = (a, min(b, nx), c, min(d, ny))
        fo = O.field_opts(s["chain"], np.float32(s["missing"]), np.float32(s["fill"]), s["masked"], s["mask_val"])
        O.interp_met_field(src, fo, s["stagger"], xlat, xlon, 1, 1, _G["sd"], s["slab"], -2, 1, 3, mask=s["mask"],
                           land_mask=s["land_mask"], water_mask=s["water_mask"],
                           landmask=_G["landmask"] if s["use_landmask"] else None, sub=patch)
    return 0
