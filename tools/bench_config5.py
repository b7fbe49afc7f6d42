#!/usr/bin/env python
"""BASELINE config 5 shape on one GPU: ERA5-like 0.25-degree source, 137 model levels of TT (mass), UU and VV
(staggered) -> 1-km 4000x4000 Lambert domain, device-resident, default sixteen_pt chain.  Prints one JSON line per
kernel choice (WPSGPU_TILE16=0 point-major packed kernel, =1 tile-major kernel) so the two can be compared where a
source cell covers ~770 destination points instead of ~68 (config 2).  python tools/bench_config5.py [--steps K]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--nlev", type=int, default=137)
    ap.add_argument("--n", type=int, default=4000)
    args = ap.parse_args()
    import torch
    import bench
    from wps_b200 import capi, synth
    dev = torch.device("cuda", 0)
    config = dict(src=(1440, 721, -0.25, 0.25), nlev=args.nlev, dom=(args.n, args.n, 1000.0, 38.5, 38.5, -97.5, 38.5, -97.5))
    table = []
    for field, st, kind in (("TT", capi.M, "smooth"), ("UU", capi.U, "wind"), ("VV", capi.V, "wind")):
        table.append(dict(field=field, nlev=args.nlev, stagger=st, chain=synth.C3D, missing=1e20, fill=0.0, masked=capi.NOT_MASKED,
                          mask=None, kind=kind))
    synth.gfs_slab_table = lambda nlev: table   # build_workload asks synth for the step's slab table
    h = capi.Handle(0, capi.PTR_DEVICE)
    w = bench.build_workload(h, torch, dev, config)
    for gid, (xlat, xlon) in w["grids"].items():
        h.metgrid_set_grid(gid, xlat, xlon, 1, 1)
    h.metgrid_set_landmask(0, w["landmask"])
    h.metgrid_set_domain(1, w["nx"], 1, w["ny"])
    descs = bench.slab_descs(capi, w)
    pts = sum(p["r"].shape[0] * p["r"].shape[1] * p["r"].shape[2] for p in w["plan"])
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    ref = None
    for mode in ("0", "1"):
        os.environ["WPSGPU_TILE16"] = mode
        for _ in range(2):
            bench.enqueue_all(h, w, descs)
        h.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            bench.enqueue_all(h, w, descs)
        ev1.record(stream)
        h.synchronize()
        ms = ev0.elapsed_time(ev1) / args.steps
        t3, p3 = h.metgrid_last_timings()
        outs = [p["r"].clone() for p in w["plan"][:1]]
        same = None
        if ref is None:
            ref = outs
        else:
            same = all(bool((a.view(torch.int32) == b.view(torch.int32)).all()) for a, b in zip(ref, outs))
        print(json.dumps({"workload": "config5 shape: 0.25deg x %d levels x (TT,UU,VV) -> %dx%d 1-km" % (args.nlev, args.n, args.n),
                          "WPSGPU_TILE16": mode, "points_per_step": pts, "ms_per_step": ms, "points_per_s": pts / (ms * 1e-3),
                          "sixteen_pt_kernel_ms": t3[1], "pack16_ms": t3[0], "slowbits_ms": t3[4],
                          "alg_GBps_sixteen_pt_kernel": 4.5 * p3[1] / (t3[1] * 1e-3) / 1e9 if t3[1] > 0 else None,
                          "tt_identical_to_point_major": same}), flush=True)


if __name__ == "__main__":
    main()
