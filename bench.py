#!/usr/bin/env python
"""bench.py -- metgrid output points/s (field x level x destination point) on the BASELINE.json configs[1]
workload: GFS 0.25 deg (1440x721, 34 isobaric levels + surface, default METGRID.TBL.ARW chains) ->
3-km CONUS Lambert 1799x1059.  One step = one time of the FILE (190 slabs, 3.6e8 output points).

  python bench.py --gpus N --steps K --warmup W           our arm (libwpsgpu.so through the C ABI)
  python bench.py --impl reference ...                    the reference's CPU path on the host cores
                                                          (oracle restatement; WPS itself cannot be built here)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_POINT = 4.5   # 4 B store + source bbox + cached coords/L + mask bit
METRIC = "metgrid output points/s (field*level*pt)"
WORKLOAD = "config2: GFS 0.25deg 1440x721 x (34 isobaric + sfc) levels, default METGRID.TBL.ARW chains -> 3-km CONUS Lambert 1799x1059; 1 step = 1 time (190 slabs)"


# Arithmetic of the sixteen_pt chains (wpsgpu_metgrid_set_precision): 1 = north_star tolerance (1e-5, strip kernel, the
# headline), 0 = bit-exact (k_metgrid_fast16).  WPSGPU_PRECISION=0 times the exact mode for A/B lines.
PRECISION = int(os.environ.get("WPSGPU_PRECISION", "1"))
K16 = "k_metgrid_strip16" if PRECISION == 1 else "k_metgrid_fast16"


def env_int(k, d):
    return int(os.environ.get(k, d))


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def step_config(config, pts):
    """The `config` object of the JSON line: identical for the device arm and the reference arm."""
    from wps_b200 import synth
    return {"workload": WORKLOAD, "slabs_per_step": synth.n_slabs(synth.gfs_slab_table(config["nlev"])), "points_per_step_per_gpu": pts,
            "l2": "inputs (792 MB of source slabs) and outputs (1.45 GB) per step exceed the 126 MB L2",
            "sharding": "one FILE time per GPU per step (field x level x time shard, no collective)",
            "inputs": "synth.gfs_step_numpy(config2): the same 190 slabs, masks and values in both arms"}


def cpu_step(config):
    """The whole step as the list of interp_met_field calls the CPU arm makes: every slab of synth.gfs_step_numpy."""
    from wps_b200 import synth
    ls, fields = synth.gfs_step_numpy(config)
    lsh = synth.add_halo(ls)
    calls = []
    for f in fields:
        e = f["e"]
        mask, lmask, wmask, mv = None, None, None, (1e20, 1e20, 1e20)
        if e["mask"] is not None:
            if e["mask"][0] == "mask":
                mask, mv = lsh, (e["mask"][1], 1e20, 1e20)
            else:
                lmask, wmask, mv = lsh, lsh, (1e20, e["mask"][1], e["mask"][2])
        for k in range(e["nlev"]):
            calls.append(dict(name=e["field"], chain=e["chain"], stagger=e["stagger"], missing=e["missing"], fill=e["fill"],
                              masked=e["masked"], mask_val=mv, mask=mask, land_mask=lmask, water_mask=wmask,
                              use_landmask=(e["stagger"] == 1), slab=f["slabs"][k]))
    return calls


_CPU = {}


def cpu_arm(config, nprocs, calls=None):
    """One step (all 190 slabs) through the oracle restatement of interp_met_field on `nprocs` host processes."""
    from oracle import oracle as O, cpu_bench
    from wps_b200 import synth
    sx, sy, dlat, dlon = config["src"]
    nx, ny, dx, tl1, tl2, slon, rlat, rlon = config["dom"]
    O.set_transc_mode(0)
    if "grids" not in _CPU:
        dom = O.map_set(O.PROJ_LC, truelat1=tl1, truelat2=tl2, stdlon=slon, lat1=rlat, lon1=rlon, knowni=(nx + 1) / 2.0,
                        knownj=(ny + 1) / 2.0, dx=dx)
        _CPU["grids"] = {1: O.get_lat_lon_fields(dom, 1, 1, nx, ny, O.M), 2: O.get_lat_lon_fields(dom, 1, 1, nx + 1, ny, O.U),
                         3: O.get_lat_lon_fields(dom, 1, 1, nx, ny + 1, O.V)}
        lam = np.deg2rad(_CPU["grids"][1][1]) % (2 * np.pi)
        _CPU["landmask"] = (synth.landsea_fn(np, lam, np.deg2rad(_CPU["grids"][1][0])) > synth.LAND_THRESHOLD).astype(np.float32)
        _CPU["calls"] = cpu_step(config)
    calls = calls if calls is not None else _CPU["calls"]
    src_kw = dict(dlat=dlat, dlon=dlon, known_x=1.0, known_y=1.0, known_lat=90.0, known_lon=0.0, earth_radius=6371229.0)
    pps, dt, pts, nprocs = cpu_bench.run(src_kw, _CPU["grids"], _CPU["landmask"], (1, nx, 1, ny), calls, nprocs, 1)
    desc = "%d slabs (the whole step of synth.gfs_step_numpy: same table, masks and values as the device arm) on the full 1799x1059 grid, %d-process patch decomposition as in the dmpar build" % (len(calls), nprocs)
    return pps, dt, pts, nprocs, desc


def run_reference(args, rank):
    from wps_b200 import synth
    if rank != 0:
        return
    config = synth.CONFIGS["config2"]
    nprocs = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        cpu_arm(config, nprocs)
    times, pts = [], 0
    for _ in range(args.steps):
        pps, dt, p, nprocs, desc = cpu_arm(config, nprocs)
        times.append(dt)
        pts = p
    ms = 1e3 * float(np.mean(times))
    value = pts / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": step_config(config, pts),
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": nprocs, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def build_workload(h, torch, dev, config, device_mode=True):
    """Synthetic inputs resident in HBM.  Returns dict with grids, slabs, outputs and the enqueue plan."""
    from wps_b200 import capi, synth
    sx, sy, dlat, dlon = config["src"]
    nx, ny, dx, tl1, tl2, slon, rlat, rlon = config["dom"]
    dom = synth.lambert_domain(nx, ny, dx, tl1, tl2, slon, rlat, rlon)
    src = capi.push_source_projection(capi.PROJ_LATLON, dlat=dlat, dlon=dlon, known_x=1.0, known_y=1.0, known_lat=90.0,
                                      known_lon=0.0, earth_radius=6371229.0)
    grids = {}
    for gid, (st, gx, gy) in {0: (capi.M, nx, ny), 1: (capi.U, nx + 1, ny), 2: (capi.V, nx, ny + 1)}.items():
        xlat = torch.empty((gy, gx), dtype=torch.float32, device=dev)
        xlon = torch.empty_like(xlat)
        h.xytoll_grid(dom, st, 1, 1, gx, gy, out=(xlat, xlon))   # get_lat_lon_fields on device
        grids[gid] = (xlat, xlon)
    h.synchronize()
    lam = torch.deg2rad(grids[0][1]) % (2 * np.pi)
    phi = torch.deg2rad(grids[0][0])
    landmask = (synth.landsea_fn(torch, lam, phi) > synth.LAND_THRESHOLD).float().contiguous()
    ls_np, fields = synth.gfs_step_numpy(config)
    ls = torch.from_numpy(ls_np).to(dev)
    lsh = synth.add_halo(ls)
    plan = []
    for f in fields:
        e = f["e"]
        gid = {capi.M: 0, capi.U: 1, capi.V: 2}[e["stagger"]]
        gy, gx = grids[gid][0].shape
        n = e["nlev"]
        slabs = torch.from_numpy(f["slabs"]).to(dev)
        masks, mv = (None, None, None), (1e20, 1e20, 1e20)
        if e["mask"] is not None:
            if e["mask"][0] == "mask":
                masks, mv = (lsh, None, None), (e["mask"][1], 1e20, 1e20)
            else:
                masks, mv = (None, lsh, lsh), (1e20, e["mask"][1], e["mask"][2])
        fo = capi.field_opts(e["chain"], np.float32(e["missing"]), np.float32(e["fill"]), e["masked"], mv)
        r = torch.empty((n, gy, gx), dtype=torch.float32, device=dev)
        npw = torch.zeros((n, gy, capi.nwords(gx)), dtype=torch.int32, device=dev)
        plan.append(dict(e=e, gid=gid, fo=fo, slabs=slabs, masks=masks, r=r, npw=npw,
                         use_landmask=(e["stagger"] == capi.M)))
    return dict(dom=dom, src=src, grids=grids, landmask=landmask, plan=plan, nx=nx, ny=ny)


def slab_descs(capi, w, host=None, raw=False):
    """wpsgpu_slab descriptors of the whole step, built once (the buffers are re-used by every step).
    raw: the slabs are big-endian records without halo, as they sit in an intermediate file."""
    out = []
    for pi, p in enumerate(w["plan"]):
        bufs = p if host is None else host[pi]
        n = bufs["slabs"].shape[0]
        arr = (capi.Slab * n)()
        for l in range(n):
            capi.Handle.slab_desc(bufs["raw"][l] if raw else bufs["slabs"][l], -2, 1, 3, p["gid"], p["e"]["stagger"], bufs["r"][l],
                                  bufs["npw"][l], masks=bufs["masks"], use_landmask=p["use_landmask"], out=arr[l],
                                  raw_big_endian=True if raw else None)
        out.append((p["fo"], arr, n))
    return out


def enqueue_all(h, w, descs):
    """The process_intermediate_fields record loop: one enqueue call per field (its levels), then ONE flush."""
    for fo, arr, n in descs:
        h.metgrid_enqueue_slabs(w["src"], fo, arr, n)
    h.metgrid_flush()


def vertical_fill_arm(h, torch, dev, config, steps):
    """SURVEY 8(f) row 1 beside the headline: fill_missing_levels' vertical interpolation over one config-2 3-d field
    (35 levels x 1799 x 1059, 30 % of the points missing at random) resident in HBM."""
    nx, ny = config["dom"][0], config["dom"][1]
    nlev = config["nlev"] + 1
    from wps_b200 import capi
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    levels = [1000 + 2900 * k for k in range(nlev)]
    r0 = torch.rand((nlev, ny, nx), device=dev, generator=g) * 100.0 + 200.0
    miss = torch.rand((nlev, ny, nx), device=dev, generator=g) < 0.3
    nw = capi.nwords(nx)
    pad = torch.zeros((nlev, ny, nw * 32), dtype=torch.bool, device=dev)
    pad[:, :, :nx] = ~miss
    weights = (2 ** torch.arange(32, device=dev, dtype=torch.int64))
    v0 = (pad.view(nlev, ny, nw, 32).to(torch.int64) * weights).sum(-1)
    v0 = torch.where(v0 >= 2 ** 31, v0 - 2 ** 32, v0).to(torch.int32).contiguous()
    r, v = r0.clone(), v0.clone()
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for it in range(steps + 1):
        r.copy_(r0); v.copy_(v0)
        torch.cuda.synchronize()
        ev0.record(stream)
        h.fill_missing_levels(levels, [r[k] for k in range(nlev)], [v[k] for k in range(nlev)])
        ev1.record(stream)
        h.synchronize()
        if it > 0:
            ms.append(ev0.elapsed_time(ev1))
    filled = int(((v != v0).sum()).item())   # words that changed (diagnostic only)
    nmiss = int(miss[1:-1].sum().item())
    t = float(np.mean(ms))
    alg = nlev * ny * nw * 4 + 12.0 * nmiss   # every valid word once + (2 reads + 1 write) per interpolated point
    return {"metric": "fill_missing_levels points/s (level x point)", "value": nlev * nx * ny / (t / 1e3), "unit": "points/s",
            "ms": t, "levels": nlev, "missing_points": nmiss, "changed_words": filled,
            "alg_GBps": alg / (t * 1e-3) / 1e9}


def geogrid_arm(h, torch, dev, config, steps):
    """geogrid reported additionally as source pixels/s.  Config-3 shape: a 30-arc-second, 21-category,
    1-byte land-use dataset in 1200x1200 tiles (raw bytes resident in HBM) accumulated onto the 3-km CONUS patch
    (+-3 halo) with calc_field's categorical path: decode, where_maps_to, quadtree credit, per-tile merge, per-point
    fallback, finish.  One step = one whole field (all tiles)."""
    from wps_b200 import capi, synth
    nx, ny, dx, tl1, tl2, slon, rlat, rlon = config["dom"]
    dom = synth.lambert_domain(nx, ny, dx, tl1, tl2, slon, rlat, rlon)
    gx, gy, ncat, tn = nx + 6, ny + 6, 21, 1200
    xlat = torch.empty((gy, gx), dtype=torch.float32, device=dev)
    xlon = torch.empty_like(xlat)
    h.xytoll_grid(dom, capi.M, -2, -2, gx, gy, out=(xlat, xlon))
    h.synchronize()
    dxdeg, klat, klon = 0.00833333, -89.99583, -179.99583
    src = capi.push_source_projection(capi.PROJ_LATLON, dlat=dxdeg, dlon=dxdeg, known_x=1.0, known_y=1.0, known_lat=klat,
                                      known_lon=klon)
    lat0, lat1 = float(xlat.min()) - 0.1, float(xlat.max()) + 0.1
    lon0, lon1 = float(xlon.min()) - 0.1, float(xlon.max()) + 0.1
    ix0, ix1 = int(np.floor((lon0 - klon) / dxdeg)) + 1, int(np.ceil((lon1 - klon) / dxdeg)) + 1
    iy0, iy1 = int(np.floor((lat0 - klat) / dxdeg)) + 1, int(np.ceil((lat1 - klat) / dxdeg)) + 1
    tiles = []
    for ty in range(tn * ((iy0 - 1) // tn) + 1, iy1 + 1, tn):
        for tx in range(tn * ((ix0 - 1) // tn) + 1, ix1 + 1, tn):
            j = torch.arange(ty, ty + tn, device=dev)[:, None]
            i = torch.arange(tx, tx + tn, device=dev)[None, :]
            cat = (((i // 7) * 31 + (j // 5) * 17) % ncat + 1).to(torch.uint8).contiguous()
            tiles.append((tx, ty, cat))
    field = torch.zeros((ncat, gy, gx), dtype=torch.float32, device=dev)
    stream = torch.cuda.ExternalStream(h.stream, device=dev)

    def one_field():
        h.geogrid_field_begin(dom, capi.M, xlat, xlon, -2, -2, 1, field)
        h.geogrid_level_begin(src, 1, capi.CATEGORICAL, "nearest_neighbor", msg_val=255.0)
        for tx, ty, cat in tiles:
            h.geogrid_add_tile_raw(cat, (1, tn, tn), 1, 0, 0, 1, tx, ty, 1, 0)
        h.geogrid_level_end()
        return h.geogrid_field_finish()

    one_field()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = h.launch_count
    ev0.record(stream)
    for _ in range(steps):
        one_field()
    ev1.record(stream)
    h.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    pixels = len(tiles) * tn * tn
    frac_sum = float(field.sum(dim=0).mean())
    launches_per_field = (h.launch_count - l0) // steps
    # HGT_M on the same patch: 30-arc-second topography, 2-byte signed big-endian tiles with tile_bdr=3.  At 3 km the
    # average_gcell condition of calc_field (:1306-1317) fails, so every point takes four_pt+average_4pt (get_point).
    bdr = 3
    ttiles = []
    for tx, ty, _ in tiles:
        j = torch.arange(ty - bdr, ty + tn + bdr, device=dev, dtype=torch.float32)[:, None]
        i = torch.arange(tx - bdr, tx + tn + bdr, device=dev, dtype=torch.float32)[None, :]
        hgt = (1500.0 + 1200.0 * torch.sin(i / 370.0) * torch.cos(j / 530.0)).round().to(torch.int16)
        be = torch.stack([(hgt >> 8).to(torch.uint8), (hgt & 255).to(torch.uint8)], dim=-1).contiguous()   # big-endian bytes
        ttiles.append((tx - bdr, ty - bdr, be))
    hfield = torch.zeros((1, gy, gx), dtype=torch.float32, device=dev)

    def one_topo():
        h.geogrid_field_begin(dom, capi.M, xlat, xlon, -2, -2, 1, hfield)
        h.geogrid_level_begin(src, 1, capi.CONTINUOUS, "four_pt+average_4pt", msg_val=-32768.0)
        for tx, ty, be in ttiles:
            h.geogrid_add_tile_raw(be, (1, tn + 2 * bdr, tn + 2 * bdr), 2, 1, 0, 1, tx, ty, 1, bdr)
        h.geogrid_level_end()
        return h.geogrid_field_finish()

    one_topo()
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(steps):
        one_topo()
    ev1.record(stream)
    h.synchronize()
    topo_ms = ev0.elapsed_time(ev1) / steps
    topo_mean = float(hfield.mean())
    return {"metric": "geogrid source pixels/s (categorical accumulation, config-3 land use)", "value": pixels / (ms / 1e3),
            "unit": "pixels/s", "ms_per_field": ms, "tiles": len(tiles), "tile": "1200x1200x1B", "pixels_per_field": pixels,
            "dest": "%dx%dx%d categories" % (gx, gy, ncat), "gpu_launches_per_field": launches_per_field,
            "alg_bytes_per_pixel": 1.0 + 4.0 * ncat * gx * gy / pixels, "mean_fraction_sum": frac_sum,
            "topo_four_pt": {"ms_per_field": topo_ms, "points_per_s": gx * gy / (topo_ms / 1e3), "tile": "1206x1206x2B signed, tile_bdr 3",
                             "pixels_per_s": len(ttiles) * (tn + 2 * bdr) ** 2 / (topo_ms / 1e3), "mean_height": topo_mean}}



def aux_kernels_arm(h, torch, dev, steps=5):
    """The other hot-path subsystems north_star lists, timed alone on config-2 / config-3 sized arrays resident in HBM:
    smooth_module (one smth-desmth_special pass and three 1-2-1 passes over HGT_M), rotate_winds (met_to_map, 35 levels of
    U / V), fractions + dominant category + landmask (21 categories).  GB/s = algorithmic bytes / device time."""
    from wps_b200 import capi, synth
    nx, ny, dx, tl1, tl2, slon, rlat, rlon = synth.CONFIGS["config2"]["dom"]
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.Generator(device=dev)
    g.manual_seed(3)

    def timed(fn, prep=None):
        ms = []
        for it in range(steps + 1):
            if prep:
                prep()
            torch.cuda.synchronize()
            ev0.record(stream)
            fn()
            ev1.record(stream)
            h.synchronize()
            if it > 0:
                ms.append(ev0.elapsed_time(ev1))
        return float(np.mean(ms))
    out = {}
    gx, gy = nx + 6, ny + 6
    hgt0 = (torch.rand((1, gy, gx), device=dev, generator=g) * 3000.0 - 100.0).contiguous()
    hgt = hgt0.clone()
    n = gx * gy
    t = timed(lambda: h.smooth(hgt, capi.SMTHDESMTH_SPECIAL, 1), lambda: hgt.copy_(hgt0))
    # one pass: read + write by the fused kernel, then the scratch is copied back into the caller's array (device-pointer mode)
    out["smth_desmth_special_1pass"] = {"ms": t, "points": n, "alg_bytes_per_point": 8.0, "GBps": 8.0 * n / (t * 1e-3) / 1e9,
                                        "launches": "1 fused pass (tile + apron of 2 in shared memory) + 1 copy back"}
    t = timed(lambda: h.smooth(hgt, capi.ONETWOONE, 3), lambda: hgt.copy_(hgt0))
    out["one_two_one_3pass"] = {"ms": t, "points": n, "alg_bytes_per_point": 24.0, "GBps": 24.0 * n / (t * 1e-3) / 1e9,
                                "launches": "3 fused passes, the last one in place"}
    nlev = 35
    dom = synth.lambert_domain(nx, ny, dx, tl1, tl2, slon, rlat, rlon)
    xlat_u = torch.empty((ny, nx + 1), device=dev); xlon_u = torch.empty_like(xlat_u)
    xlat_v = torch.empty((ny + 1, nx), device=dev); xlon_v = torch.empty_like(xlat_v)
    h.xytoll_grid(dom, capi.U, 1, 1, nx + 1, ny, out=(xlat_u, xlon_u))
    h.xytoll_grid(dom, capi.V, 1, 1, nx, ny + 1, out=(xlat_v, xlon_v))
    u = (torch.rand((nlev, ny, nx + 1), device=dev, generator=g) * 40.0 - 20.0).contiguous()
    v = (torch.rand((nlev, ny + 1, nx), device=dev, generator=g) * 40.0 - 20.0).contiguous()
    um = torch.full((nlev, ny, capi.nwords(nx + 1)), -1, dtype=torch.int32, device=dev)
    vm = torch.full((nlev, ny + 1, capi.nwords(nx)), -1, dtype=torch.int32, device=dev)
    t = timed(lambda: h.rotate_winds(-1, dom, u, um, v, vm, xlon_u, xlon_v))
    npt = u.numel() + v.numel()
    # per output: own value (4 B) + four neighbours of the other component (mostly cached: ~4 B) + store (4 B)
    out["rotate_winds_met_to_map"] = {"ms": t, "points": npt, "levels": nlev, "alg_bytes_per_point": 12.0, "GBps": 12.0 * npt / (t * 1e-3) / 1e9}
    ncat = 21
    cnt0 = torch.randint(0, 9, (ncat, gy, gx), device=dev, generator=g).float().contiguous()
    cnt = cnt0.clone()
    lm = torch.empty((gy, gx), device=dev); dm = torch.empty((gy, gx), device=dev)
    t = timed(lambda: h.fractions_dominant(cnt, 1, 1e20, [17], True, lm, dm), lambda: cnt.copy_(cnt0))
    out["fractions_dominant_landmask"] = {"ms": t, "points": n, "categories": ncat, "alg_bytes_per_point": 8.0 * ncat + 8.0,
                                          "GBps": (8.0 * ncat + 8.0) * n / (t * 1e-3) / 1e9}
    return out


def mpas_arm(h, torch, dev, steps=5):
    """an MPAS source (synthetic 40962-cell Voronoi mesh, 55 levels) remapped onto the config-2 mass grid with
    process_mpas_fields' per-level planes written directly (wpsgpu_mpas_remap_planes).  Algorithmic bytes per output value:
    4 B store + (3 indices + 3 weights) = 24 B per point amortised over the levels; the source columns are L2 hits."""
    import time
    from wps_b200 import capi, synth
    nx, ny, dx, tl1, tl2, slon, rlat, rlon = synth.CONFIGS["config2"]["dom"]
    mesh = synth.mpas_mesh(40962, seed=6)
    dom = synth.lambert_domain(nx, ny, dx, tl1, tl2, slon, rlat, rlon)
    xlat = torch.empty((ny, nx), device=dev); xlon = torch.empty_like(xlat)
    h.xytoll_grid(dom, capi.M, 1, 1, nx, ny, out=(xlat, xlon))
    d2r = float(np.float32(np.arcsin(np.float32(1.0)) / np.float32(90.0)))
    lat, lon = (xlat * d2r).contiguous(), (xlon * d2r).contiguous()
    h.mpas_set_mesh(mesh)
    torch.cuda.synchronize()
    n0 = h.launch_count
    t0 = time.perf_counter()
    h.mpas_remap_setup(0, lat, lon)
    h.synchronize()
    setup_s = time.perf_counter() - t0
    setup_launches = h.launch_count - n0
    nlev = 55
    src = torch.from_numpy(synth.mpas_field(mesh, nlev, seed=1)).to(dev)
    planes = [torch.empty((ny, nx), device=dev) for _ in range(nlev)]
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for it in range(steps + 2):
        torch.cuda.synchronize()
        ev0.record(stream)
        h.mpas_remap_planes(0, capi.MPAS_CELLS, src, planes)
        ev1.record(stream)
        h.synchronize()
        if it > 1:
            ms.append(ev0.elapsed_time(ev1))
    t = float(np.mean(ms))
    npt = nx * ny * nlev
    bpp = 4.0 + 24.0 / nlev
    return {"metric": "MPAS remap output values/s (level x target point), planes written in the met_em layout", "value": npt / (t * 1e-3),
            "unit": "points/s", "ms": t, "points": npt, "levels": nlev, "mesh_cells": mesh["nCells"], "target": "%dx%d" % (nx, ny),
            "alg_bytes_per_point": bpp, "GBps": bpp * npt / (t * 1e-3) / 1e9, "launches_per_field": 1,
            "setup": {"seconds": setup_s, "launches": int(setup_launches),
                      "what": "remap_info_setup for 1.9 M target points: nearest cell / vertex / edge, four Wachspress tables, search_for_cells; "
                              "passes of the sequential-scan recurrence until it settles"}}


def config4_arm(h, torch, dev, steps=5):
    """BASELINE configs[3]: 27 / 9 / 3-km nests (compute_nest_locations), Lambert-grid source (NAM-218-like, 614 x 428),
    per nest TT / UU / VV x 8 levels + SST (masked land) + soil temperature (masked water, search) + SKINTEMP (masked=both),
    ONE flush per nest, then map_to_met + met_to_map on the winds.  Device-resident; small domains: launch-latency bound."""
    from wps_b200 import capi, synth
    cfg = synth.CONFIGS["config4"]
    nests = [(p, r, (ij if isinstance(ij, tuple) else (1, 1))[0], (ij if isinstance(ij, tuple) else (1, 1))[1], ewe, esn)
             for (p, r, ij, ewe, esn) in cfg["nests"]]
    nx1, ny1, dx, tl1, tl2, slon, rlat, rlon = cfg["dom"]
    projs = capi.compute_nest_locations(capi.PROJ_LC, nests, truelat1=tl1, truelat2=tl2, stdlon=slon, lat1=rlat, lon1=rlon,
                                        knowni=nx1 / 2.0, knownj=ny1 / 2.0, dx=dx)
    snx, sny, nlev = 614, 428, cfg["nlev"]
    src = capi.push_source_projection(capi.PROJ_LC, stand_lon=-95.0, truelat1=25.0, truelat2=25.0, dxkm=12190.58, dykm=12190.58,
                                      known_x=1.0, known_y=1.0, known_lat=12.19, known_lon=-133.459)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    jj = torch.arange(sny, device=dev, dtype=torch.float32)[:, None]
    ii = torch.arange(snx, device=dev, dtype=torch.float32)[None, :]
    ls = ((torch.sin(ii / 23.0) + torch.cos(jj / 17.0)) > -0.2).float().contiguous()
    tt = (285.0 + 8.0 * torch.sin(ii / 60.0) * torch.cos(jj / 45.0))[None].repeat(nlev, 1, 1).contiguous()
    uu = (12.0 * torch.sin(ii / 70.0) * torch.cos(jj / 50.0))[None].repeat(nlev, 1, 1).contiguous()
    sst = (290.0 + 6.0 * torch.cos(jj / 80.0) + 0.0 * ii).contiguous()
    sst[ls == 1] = -1e30
    soil = (280.0 + 5.0 * torch.sin(ii / 40.0) + 0.0 * jj).contiguous()
    soil[ls == 0] = -1e30
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms, total_pts, launches = 0.0, 0, 0
    per_nest = []
    for dom_id, (pg, nest) in enumerate(zip(projs, nests), start=1):
        nx, ny = nest[4] - 1, nest[5] - 1
        grids = {}
        for gid, (st, gx, gy) in {0: (capi.M, nx, ny), 1: (capi.U, nx + 1, ny), 2: (capi.V, nx, ny + 1)}.items():
            xlat = torch.empty((gy, gx), dtype=torch.float32, device=dev); xlon = torch.empty_like(xlat)
            h.xytoll_grid(pg, st, 1, 1, gx, gy, out=(xlat, xlon))
            h.metgrid_set_grid(gid, xlat, xlon, 1, 1)
            grids[gid] = (xlat, xlon)
        h.metgrid_set_domain(1, nx, 1, ny)
        lm = ((torch.sin(grids[0][1] * 0.9) + torch.cos(grids[0][0] * 1.3)) > -0.3).float().contiguous()
        h.metgrid_set_landmask(0, lm)
        fields = [(tt, 0, capi.M, synth.C3D, 1e20, 0.0, -2.0, (None, None, None), (1e20, 1e20, 1e20), True),
                  (uu, 1, capi.U, synth.C3D, 1e20, 0.0, -2.0, (None, None, None), (1e20, 1e20, 1e20), False),
                  (uu, 2, capi.V, synth.C3D, 1e20, 0.0, -2.0, (None, None, None), (1e20, 1e20, 1e20), False),
                  (sst[None], 0, capi.M, "sixteen_pt+four_pt+wt_average_4pt+search", -1e30, 0.0, 1.0, (ls, None, None), (1.0, 1e20, 1e20), True),
                  (soil[None], 0, capi.M, synth.SOIL, -1e30, 285.0, 0.0, (ls, None, None), (0.0, 1e20, 1e20), True),
                  (tt[:1], 0, capi.M, synth.SOIL, 1e20, 0.0, -1.0, (None, ls, ls), (1e20, 1.0, 0.0), True)]
        plan, pts = [], 0
        for slabs, gid, st, chain, msg, fill, masked, masks, mv, use_lm in fields:
            gy, gx = grids[gid][0].shape
            n = slabs.shape[0]
            r = torch.zeros((n, gy, gx), dtype=torch.float32, device=dev)
            npw = torch.zeros((n, gy, capi.nwords(gx)), dtype=torch.int32, device=dev)
            arr = (capi.Slab * n)()
            for k in range(n):
                capi.Handle.slab_desc(slabs[k], 1, 1, 0, gid, st, r[k], npw[k], masks=masks, use_landmask=use_lm, out=arr[k])
            plan.append((capi.field_opts(chain, np.float32(msg), np.float32(fill), masked, mv), arr, n, r, npw))
            pts += n * gy * gx

        def one():
            for fo, arr, n, _, _ in plan:
                h.metgrid_enqueue_slabs(src, fo, arr, n)
            h.metgrid_flush()
            h.rotate_winds(1, src, plan[1][3], plan[1][4], plan[2][3], plan[2][4], grids[1][1], grids[2][1])
            h.rotate_winds(-1, pg, plan[1][3], plan[1][4], plan[2][3], plan[2][4], grids[1][1], grids[2][1])
        one(); one()
        torch.cuda.synchronize()
        l0 = h.launch_count
        ev0.record(stream)
        for _ in range(steps):
            one()
        ev1.record(stream)
        h.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        per_nest.append({"domain": dom_id, "grid": "%dx%d" % (nx, ny), "ms": ms, "points": pts})
        total_ms += ms; total_pts += pts; launches += (h.launch_count - l0) // steps
    return {"metric": "metgrid output points/s, config 4 (three nests, one flush + wind rotations per nest)", "value": total_pts / (total_ms / 1e3),
            "unit": "points/s", "ms_total": total_ms, "points": total_pts, "gpu_launches": launches, "per_nest": per_nest,
            "note": "25-28 k points per nest and level: launch-latency bound, not a roofline case"}


def config5_plan(capi, synth, config, rank, world):
    """One time of an ERA5-shaped FILE: TT, QV, UU, VV on `nlev` model levels; the slabs of the step are sharded by level
    (shard.shard_slabs keeps UU and VV of a level on the same rank) -> this rank's list of (field, level, stagger, kind)."""
    from wps_b200 import shard
    fields = (("TT", capi.M, "smooth"), ("QV", capi.M, "smooth"), ("UU", capi.U, "wind"), ("VV", capi.V, "wind"))
    keys, meta = [], []
    for f, st, kind in fields:
        for l in range(config["nlev"]):
            keys.append((f, l))
            meta.append((f, l, st, kind))
    mine = shard.shard_slabs(keys, world)[rank]
    return [meta[i] for i in mine], len(keys)


def run_config5(args, rank, world, local_rank):
    """BASELINE configs[4]: 137 model levels -> 1-km 4000 x 4000, ONE time step sharded field x level over the GPUs (strong
    scaling: total work fixed).  Device-timed (HBM-resident) and end to end with pinned host buffers."""
    import torch
    import torch.distributed as dist
    from wps_b200 import capi, synth
    affinity = bind_to_gpu_numa_node(local_rank) if world > 1 else "unchanged (single process)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    config = synth.CONFIGS["config5"]
    sx, sy, dlat, dlon = config["src"]
    nx, ny, dx, tl1, tl2, slon, rlat, rlon = config["dom"]
    mine, total_slabs = config5_plan(capi, synth, config, rank, world)
    h = capi.Handle(local_rank, capi.PTR_DEVICE)
    h.metgrid_set_precision(PRECISION)
    dom = synth.lambert_domain(nx, ny, dx, tl1, tl2, slon, rlat, rlon)
    src = capi.push_source_projection(capi.PROJ_LATLON, dlat=dlat, dlon=dlon, known_x=1.0, known_y=1.0, known_lat=90.0, known_lon=0.0,
                                      earth_radius=6371229.0)
    grids, gid_of = {}, {capi.M: 0, capi.U: 1, capi.V: 2}
    for st, (gx, gy) in {capi.M: (nx, ny), capi.U: (nx + 1, ny), capi.V: (nx, ny + 1)}.items():
        xlat = torch.empty((gy, gx), dtype=torch.float32, device=dev)
        xlon = torch.empty_like(xlat)
        h.xytoll_grid(dom, st, 1, 1, gx, gy, out=(xlat, xlon))
        h.metgrid_set_grid(gid_of[st], xlat, xlon, 1, 1)
        grids[st] = (gx, gy)
    h.metgrid_set_domain(1, nx, 1, ny)
    fo = capi.field_opts(synth.C3D, np.float32(1e20), np.float32(0.0), capi.NOT_MASKED, (1e20, 1e20, 1e20))
    # this rank's slabs, grouped by field so that the levels of a field are enqueued together
    by_field = {}
    for f, l, st, kind in mine:
        by_field.setdefault((f, st, kind), []).append(l)
    plan, pts = [], 0
    for (f, st, kind), levels in by_field.items():
        gx, gy = grids[st]
        slabs = torch.from_numpy(np.stack([synth.add_halo(synth.slab_np(sx, sy, 7000 + 1000 * gid_of[st] + 37 * len(f) + l, kind)) for l in levels])).to(dev)
        r = torch.empty((len(levels), gy, gx), dtype=torch.float32, device=dev)
        npw = torch.zeros((len(levels), gy, capi.nwords(gx)), dtype=torch.int32, device=dev)
        arr = (capi.Slab * len(levels))()
        for k in range(len(levels)):
            capi.Handle.slab_desc(slabs[k], -2, 1, 3, gid_of[st], st, r[k], npw[k], out=arr[k])
        plan.append(dict(field=f, st=st, slabs=slabs, r=r, npw=npw, arr=arr, n=len(levels)))
        pts += len(levels) * gx * gy

    def step(hd, pl):
        for p in pl:
            hd.metgrid_enqueue_slabs(src, fo, p["arr"], p["n"])
        hd.metgrid_flush()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(h.stream, device=dev)
    for _ in range(max(args.warmup, 3)):
        step(h, plan)
    h.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = h.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step(h, plan)
    ev1.record(stream)
    h.synchronize()
    barrier()
    launches = h.launch_count - l0
    kms = []
    h.metgrid_set_timing(True)          # the library's event brackets are a measurement aid: off in the timed region above
    for _ in range(min(args.steps, 3)):
        step(h, plan)
        ms, kp = h.metgrid_last_kernel()
        kms.append(ms)
        t3, p3 = h.metgrid_last_timings()
    h.metgrid_set_timing(False)
    sampler.stop_flag = True
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    tp = torch.tensor([float(pts)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tp, op=dist.ReduceOp.SUM)
    ms_per_step = float(t.item()) / args.steps
    total_pts = float(tp.item())
    value = total_pts / (ms_per_step / 1e3)
    # ---- e2e: pinned host buffers through the C ABI; host memory for at most E2E_CHUNK slabs is re-used chunk after chunk
    # (every slab of the rank is still interpolated and copied back in every step)
    e2e_value, e2e_ms, h2d, d2h = None, None, 0, 0
    if not args.no_e2e:
        chunk = env_int("WPSGPU_E2E_CHUNK", 48)
        hh = capi.Handle(local_rank, capi.PTR_HOST)
        hh.metgrid_set_precision(PRECISION)
        for st, (gx, gy) in grids.items():
            xlat = np.empty((gy, gx), np.float32); xlon = np.empty_like(xlat)
            hh.xytoll_grid(dom, st, 1, 1, gx, gy, out=(xlat, xlon))
            hh.metgrid_set_grid(gid_of[st], xlat, xlon, 1, 1)
        hh.metgrid_set_domain(1, nx, 1, ny)
        hplan = []
        for p in plan:
            gx, gy = grids[p["st"]]
            nbuf = min(chunk, p["n"])
            hs = p["slabs"].cpu().pin_memory()
            hr = torch.empty((nbuf, gy, gx), dtype=torch.float32).pin_memory()
            hn = torch.zeros((nbuf, gy, capi.nwords(gx)), dtype=torch.int32).pin_memory()
            chunks = []
            for c0 in range(0, p["n"], nbuf):
                m = min(nbuf, p["n"] - c0)
                arr = (capi.Slab * m)()
                for k in range(m):
                    capi.Handle.slab_desc(hs[c0 + k], -2, 1, 3, gid_of[p["st"]], p["st"], hr[k], hn[k], out=arr[k])
                chunks.append((arr, m))
            hplan.append((chunks, hs, hr, hn))

        def host_step():
            for chunks, _, _, _ in hplan:
                for arr, m in chunks:
                    hh.metgrid_enqueue_slabs(src, fo, arr, m)
                    hh.metgrid_flush()
        host_step()
        barrier()
        b0 = hh.metgrid_copy_bytes()
        e2e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        b1 = hh.metgrid_copy_bytes()
        h2d, d2h = (b1[0] - b0[0]) // e2e_steps, (b1[1] - b0[1]) // e2e_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms = float(te.item()) * 1e3
        e2e_value = total_pts / float(te.item())
        hh.close()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        kms_avg = float(np.mean(kms))
        achieved = ALG_BYTES_PER_POINT * kp / (kms_avg * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "config5: ERA5-shaped 0.25deg 1440x721 x 137 model levels x (TT, QV, UU, VV) -> 1-km Lambert 4000x4000; 1 step = 1 time (%d slabs, %.3g points)" % (total_slabs, total_pts),
                           "sharding": "the step's slabs by level over the GPUs (wps_b200/shard.py: UU and VV of a level on the same rank), no collective",
                           "slabs_this_rank": len(mine), "points_per_step_total": total_pts,
                           "l2": "outputs (%.1f GB on rank 0) exceed the 126 MB L2" % (pts * 4 / 1e9),
                           "precision": "tolerance (sixteen_pt <= 1e-5, north_star)" if PRECISION == 1 else "exact (bit-identical sixteen_pt)"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                             "kernel": K16, "kernel_ms": kms_avg, "alg_bytes_per_point": ALG_BYTES_PER_POINT, "kernel_points": kp,
                             "breakdown_ms": {"k_pack_win": t3[0], K16: t3[1], "k_metgrid_slowlist": t3[4], "flush_first_to_last_kernel": t3[6]},
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)", "rank": 0},
                "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                        "note": "per rank; host buffers of at most WPSGPU_E2E_CHUNK slabs per field are re-used chunk after chunk"},
                "gpu_launches": int(launches), "cpu_affinity": affinity, "clocks": sampler.result()}
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.destroy_process_group()


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run (and first-touch the pinned host buffers) on the CPUs next to that GPU, so that the
    host<->device copies of the e2e leg do not cross the socket interconnect.  Returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hdl = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "cpus %d-%d (%d)" % (allowed[0], allowed[-1], len(allowed))
    except Exception as e:   # affinity is an optimisation only
        return "unchanged (%s)" % type(e).__name__
    return "unchanged"


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from wps_b200 import capi, synth
    affinity = bind_to_gpu_numa_node(local_rank) if world > 1 else "unchanged (single process)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    config = synth.CONFIGS["config2"]
    h = capi.Handle(local_rank, capi.PTR_DEVICE)
    h.metgrid_set_precision(PRECISION)
    w = build_workload(h, torch, dev, config)
    nx, ny = w["nx"], w["ny"]
    for gid, (xlat, xlon) in w["grids"].items():
        h.metgrid_set_grid(gid, xlat, xlon, 1, 1)
    h.metgrid_set_landmask(0, w["landmask"])
    h.metgrid_set_domain(1, nx, 1, ny)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(h.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    descs = slab_descs(capi, w)
    pts = sum(p["r"].shape[0] * p["r"].shape[1] * p["r"].shape[2] for p in w["plan"])
    for _ in range(max(args.warmup, 3)):
        enqueue_all(h, w, descs)
    h.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = h.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    ev0.record(stream)
    th0 = time.perf_counter()
    for _ in range(args.steps):
        enqueue_all(h, w, descs)
        kms.append(None)
    host_issue_ms = (time.perf_counter() - th0) * 1e3 / args.steps   # host time to queue one step (no sync)
    ev1.record(stream)
    h.synchronize()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = h.launch_count - l0
    # dominant-kernel time: re-run K steps reading the library's own CUDA events around k_metgrid_interp
    kms = []
    h.metgrid_set_timing(True)          # the library's event brackets are a measurement aid: off in the timed region above
    for _ in range(args.steps):
        enqueue_all(h, w, descs)
        ms, kp = h.metgrid_last_kernel()
        kms.append(ms)
        t3, p3 = h.metgrid_last_timings()
    h.metgrid_set_timing(False)
    sampler.stop_flag = True
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * pts / (ms_per_step / 1e3)

    # ---- e2e: host buffers in pinned memory through the C ABI (H2D + kernels + D2H inside the timed region) ----
    e2e_value, e2e_s, h2d, d2h, e2e_raw_s, host_input_bytes = None, 0.0, 0, 0, None, 0
    if not args.no_e2e:
        hh = capi.Handle(local_rank, capi.PTR_HOST)
        hh.metgrid_set_precision(PRECISION)
        for gid, (xlat, xlon) in w["grids"].items():
            hh.metgrid_set_grid(gid, xlat.cpu().numpy(), xlon.cpu().numpy(), 1, 1)
        hh.metgrid_set_landmask(0, w["landmask"].cpu().numpy())
        hh.metgrid_set_domain(1, nx, 1, ny)
        host, h2d, d2h = [], 0, 0
        seen = set()
        for p in w["plan"]:
            hs = p["slabs"].cpu().pin_memory()
            hm = tuple(None if m is None else m.cpu().pin_memory() for m in p["masks"])
            hr = torch.empty(p["r"].shape, dtype=torch.float32).pin_memory()
            hn = torch.zeros(p["npw"].shape, dtype=torch.int32).pin_memory()
            # the same records the way read_met_module.F finds them in the file: big-endian words, no halo columns
            raw = torch.from_numpy(hs[:, :, 3:-3].contiguous().numpy().astype(">f4").view(np.float32)).pin_memory()
            host.append(dict(slabs=hs, masks=hm, r=hr, npw=hn, raw=raw))
            h2d += hs.numel() * 4
            for m in hm:
                if m is not None and "mask" not in seen:
                    h2d += m.numel() * 4
                    seen.add("mask")
            d2h += hr.numel() * 4 + hn.numel() * 4
        e2e_steps = max(1, args.steps)
        hdescs = slab_descs(capi, w, host)
        enqueue_all(hh, w, hdescs)  # warm-up (arena allocation, coordinate cache)
        barrier()
        b0 = hh.metgrid_copy_bytes()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            enqueue_all(hh, w, hdescs)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        b1 = hh.metgrid_copy_bytes()
        # bytes the library actually moved (chains without `search` only transfer the part of the slab they can touch);
        # host_input_bytes_per_step is the size of the input arrays handed over
        host_input_bytes = h2d
        h2d, d2h = (b1[0] - b0[0]) // e2e_steps, (b1[1] - b0[1]) // e2e_steps
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = world * pts / float(t.item())
        # same, from raw records: byte order and halo are handled on the device
        rdescs = slab_descs(capi, w, host, raw=True)
        enqueue_all(hh, w, rdescs)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            enqueue_all(hh, w, rdescs)
        torch.cuda.synchronize()
        e2e_raw_s = (time.perf_counter() - t0) / e2e_steps
        hh.close()

    geo, vfill, aux, cfg4, mp = None, None, None, None, None
    if rank == 0 and not args.no_geogrid:
        geo = geogrid_arm(h, torch, dev, config, 3)
        vfill = vertical_fill_arm(h, torch, dev, config, 3)
        aux = aux_kernels_arm(h, torch, dev)
        cfg4 = config4_arm(h, torch, dev)
        mp = mpas_arm(h, torch, dev)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None   # DRAM bytes of the dominant kernel's launches in one step, from the committed ncu capture
        try:
            # written by tools/ncu_traffic.py from an ncu capture of this very command; only accepted for the kernel and the
            # amount of work it was measured on
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_%s_traffic.json" % K16)))
            if p3[1] >= p3[2] and tj["kernel"] == K16 and abs(tj["kernel_points"] - kp) < 0.01 * kp:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        kms_avg = float(np.mean(kms))
        achieved = ALG_BYTES_PER_POINT * kp / (kms_avg * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": dict(step_config(config, pts), precision="tolerance (sixteen_pt <= 1e-5, north_star)" if PRECISION == 1 else "exact (bit-identical sixteen_pt)"),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "traffic_unit": "bytes per step (all launches of the kernel), ncu dram read+write; algorithmic = %d" % int(ALG_BYTES_PER_POINT * kp),
                             "kernel": K16 if p3[1] >= p3[2] else "k_metgrid_interp", "kernel_ms": kms_avg,
                             "kernel_ms_source": "CUDA events recorded by the library around the kernel on its own stream, averaged over a second run of the same K steps (wpsgpu_metgrid_set_timing on; off in the timed region)",
                             "alg_bytes_per_point": ALG_BYTES_PER_POINT, "kernel_points": kp,
                             "breakdown_ms": {("k_pack_win" if PRECISION == 1 else "k_pack16"): t3[0], K16: t3[1], "k_metgrid_fast4": t3[3], "k_metgrid_interp(generic)": t3[2], "k_metgrid_slowlist": t3[4], "k_metgrid_search(+literal)": t3[5], "flush_first_to_last_kernel": t3[6]},
                             "breakdown_points": {"fast16": p3[1], "fast4": p3[3], "generic": p3[2]}, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                             # sixteen_pt sits at the machine's balance point: the FP32 view of the same kernel time
                             "fp32": None if PRECISION != 1 else {"fma_per_point": 30, "achieved_fma_per_s": 30.0 * kp / (kms_avg * 1e-3), "peak_fma_per_s": 148 * 128 * 1.965e9,
                                                                  "frac": 30.0 * kp / (kms_avg * 1e-3) / (148 * 128 * 1.965e9),
                                                                  "peak_source": "148 SMs x 128 FP32 lanes x 1.965 GHz (nominal)"}},
                "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3, "host_input_bytes_per_step": host_input_bytes,
                        "raw_record_ms_per_step": None if e2e_raw_s is None else e2e_raw_s * 1e3},
                "gpu_launches": int(launches), "host_issue_ms_per_step": host_issue_ms, "cpu_affinity": affinity, "clocks": sampler.result()}
        if geo is not None:
            line["geogrid"] = geo
            line["vertical_fill"] = vfill
            line["aux_kernels"] = aux
            line["config4"] = cfg4
            mp["frac_of_hbm_peak"] = mp["GBps"] / peak
            line["mpas_remap"] = mp
            for k in aux.values():
                k["frac_of_hbm_peak"] = k["GBps"] / peak
        if world == 1 and not args.no_cpu:
            pps, dt, cpts, nprocs, desc = cpu_arm(config, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": pps, "unit": "points/s", "cores": nprocs, "kind": "port", "sample": desc, "seconds": dt}
            # also the single-thread figure, on the first two slabs of the same step
            pps1, dt1, _, _, _ = cpu_arm(config, 1, calls=_CPU["calls"][:2])
            line["cpu_baseline"]["single_thread"] = {"value": pps1, "unit": "points/s", "cores": 1, "seconds": dt1,
                                                     "sample": "the first 2 TT slabs of the step (sixteen_pt chain) on the full grid, one process"}
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config5"],
                    help="config2 (default, the headline: one FILE time per GPU, weak scaling) or config5 (one 137-level time step sharded by level, strong scaling)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-geogrid", action="store_true", help="skip the additional geogrid source-pixels/s measurement")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs only)")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
    elif args.workload == "config5":
        run_config5(args, rank, world, local_rank)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
