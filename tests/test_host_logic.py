"""CPU: host-side mirrors (METGRID.TBL parser, intermediate format, sharding incl. a world_size-2 gloo run)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TBL = """
# sample in METGRID.TBL syntax
========================================
name=TT
        interp_option=sixteen_pt+four_pt+average_4pt
        fill_missing=0.
========================================
name=UU ; interp_option=sixteen_pt+four_pt+average_4pt ; is_u_field=yes ; output_stagger=U ; fill_missing=0.
========================================
name=ST000010
        interp_option=sixteen_pt+four_pt+wt_average_4pt+wt_average_16pt+search
        masked=water
        interp_mask=LANDSEA(0)
        missing_value=-1.E30   # comment
        fill_missing=285.
========================================
name=SKINTEMP ; masked=both ; interp_land_mask = LANDSEA(1) ; interp_water_mask = LANDSEA(0)
========================================
name=T ; output_name=TT
========================================
name=SST ; interp_mask=LANDSEA(>0.5) ; from_input=SSTFILE
========================================
"""


def test_read_interp_table():
    from wps_b200 import capi
    from wps_b200.interp_option import find_entry, read_interp_table
    t = read_interp_table(TBL)
    assert len(t) == 7 and t[-1].fieldname == " " and t[-1].interp_method == "nearest_neighbor"
    tt = t[find_entry(t, "TT", "FILE")]
    assert tt.interp_method == "sixteen_pt+four_pt+average_4pt" and tt.fill_missing == 0.0 and tt.missing_value == np.float32(1e20) or tt.missing_value == 1e20
    uu = t[find_entry(t, "UU", "FILE")]
    assert uu.output_stagger == capi.U and uu.is_u_field
    st = t[find_entry(t, "ST000010", "FILE")]
    assert st.masked == capi.MASKED_WATER and st.interp_mask == "LANDSEA" and st.interp_mask_val == 0.0
    assert st.missing_value == -1e30 and st.fill_missing == 285.0 and st.interp_mask_relational == " "
    sk = t[find_entry(t, "SKINTEMP", "FILE")]
    assert sk.masked == capi.MASKED_BOTH and sk.interp_land_mask_val == 1.0 and sk.interp_water_mask_val == 0.0
    assert t[find_entry(t, "T", "FILE")].output_name == "TT"
    # from_input: the SST entry only applies to sources whose name contains SSTFILE; otherwise the default entry
    assert find_entry(t, "SST", "FILE") == len(t) - 1
    sst = t[find_entry(t, "SST", "./SSTFILE")]
    assert sst.interp_mask_relational == ">" and sst.interp_mask_val == 0.5
    assert find_entry(t, "NOPE", "FILE") == len(t) - 1
    fo = st.field_opts()
    assert list(fo.ops[:6]) == [1, 2, 6, 7, 8, 0] and fo.opts[4] == 1200


def test_reference_table_if_present():
    from wps_b200.interp_option import find_entry, read_interp_table
    p = "/root/reference/metgrid/METGRID.TBL.ARW"
    if not os.path.exists(p):
        pytest.skip("reference tree not present")
    t = read_interp_table(open(p).read())
    assert t[find_entry(t, "TT", "FILE")].interp_method == "sixteen_pt+four_pt+average_4pt"
    assert t[find_entry(t, "SM000010", "FILE")].interp_mask == "LANDSEA"
    assert t[find_entry(t, "SEAICE", "FILE")].masked == 1.0


def test_intermediate_roundtrip(tmp_path):
    from wps_b200.intermediate import MetData, read_records, write_records
    rng = np.random.default_rng(0)
    recs = [MetData("TT", 85000.0, rng.uniform(200, 300, (19, 36)).astype(np.float32), startlat=90.0, startlon=0.0, deltalat=-10.0, deltalon=10.0),
            MetData("HGT", 50000.0, rng.uniform(0, 6000, (19, 36)).astype(np.float32), startlat=90.0, startlon=0.0, deltalat=-10.0, deltalon=10.0),
            MetData("UU", 200100.0, rng.uniform(-9, 9, (11, 13)).astype(np.float32), iproj=3, startlat=20.0, startlon=-110.0, dx=40.0, dy=40.0,
                    xlonc=-98.0, truelat1=30.0, truelat2=60.0, is_wind_grid_rel=True)]
    p = str(tmp_path / "FILE:2026-10-19_00")
    write_records(p, recs)
    raw = open(p, "rb").read()
    assert raw[:12] == b"\x00\x00\x00\x04\x00\x00\x00\x05\x00\x00\x00\x04"      # big-endian record marker + version 5
    back = read_records(p)
    assert [r.field for r in back] == ["TT", "GHT", "UU"]                       # HGT renamed on read
    for a, b in zip(recs, back):
        assert np.array_equal(a.slab, b.slab) and a.xlvl == b.xlvl and a.iproj == b.iproj
    assert back[2].is_wind_grid_rel and back[2].truelat2 == 60.0 and abs(back[2].earth_radius - 6371.229) < 1e-3


def test_shard_slabs_keeps_wind_pairs_together():
    from wps_b200.shard import shard_slabs, shard_times
    keys = [(f, l) for f in ("TT", "RH", "UU", "VV", "GHT") for l in range(34)] + [("PSFC", 200100), ("SKINTEMP", 200100)]
    for world in (1, 2, 4, 8):
        sh = shard_slabs(keys, world)
        allidx = sorted(i for v in sh.values() for i in v)
        assert allidx == list(range(len(keys)))
        where = {keys[i]: r for r, v in sh.items() for i in v}
        for l in range(34):
            assert where[("UU", l)] == where[("VV", l)]
        sizes = [len(v) for v in sh.values()]
        assert max(sizes) - min(sizes) <= 2
    assert shard_times(24, 1, 8) == [1, 9, 17]


def test_sharded_run_gloo_world2(tmp_path):
    """N>1 host path on CPU: two gloo ranks shard the slab list, each computes a checksum of 'its' work and the
    rank-0 gather reproduces the single-process result (no data-path collective is needed, only the final gather)."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import numpy as np, torch, torch.distributed as dist
        from wps_b200.shard import shard_slabs
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        keys = [(f, l) for f in ("TT", "UU", "VV") for l in range(10)]
        mine = shard_slabs(keys, world)[rank]
        vals = torch.zeros(len(keys), dtype=torch.float64)
        for i in mine:
            vals[i] = (i + 1) * 1.5      # stands for the slab's output checksum
        dist.all_reduce(vals)
        t = torch.tensor([float(len(mine))], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            assert torch.equal(vals, torch.arange(1, len(keys) + 1, dtype=torch.float64) * 1.5)
            print("OK", int(t.item()))
        dist.destroy_process_group()
    """ % ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK 15" in out.stdout


def test_patch_decomposition_matches_rsl_rule():
    """parallel_get_tile_dims / task_for_point (parallel_module.F:103-148, 218-267): hand-derived cases and the
    partition property (every point owned exactly once, patches are rectangles)."""
    from wps_b200.shard import _task_1d, patch_bounds, proc_grid
    assert [proc_grid(n) for n in (1, 2, 3, 4, 6, 8, 12)] == [(1, 1), (1, 2), (1, 3), (2, 2), (2, 3), (2, 4), (3, 4)]
    # idim=10 over 3 tasks: rem=1 -> a=0, b=6: sizes 3,3,4 (the remainder goes to the tail when rem/2 == 0)
    assert [_task_1d(i, 1, 10, 3) for i in range(1, 11)] == [0, 0, 0, 1, 1, 1, 2, 2, 2, 2]
    # idim=11 over 4 tasks: rem=3 -> a=1*3=3, b=3+1*2=5: sizes 3,2,3,3
    assert [_task_1d(i, 1, 11, 4) for i in range(1, 12)] == [0, 0, 0, 1, 1, 2, 2, 2, 3, 3, 3]
    for (idim, jdim, n) in [(1799, 1059, 8), (101, 101, 6), (74, 61, 4), (10, 7, 3)]:
        owner = np.zeros((jdim, idim), np.int32) - 1
        for r in range(n):
            x0, x1, y0, y1 = patch_bounds(idim, jdim, n, r)
            assert (owner[y0 - 1:y1, x0 - 1:x1] == -1).all()
            owner[y0 - 1:y1, x0 - 1:x1] = r
        assert (owner >= 0).all()


def test_patch_decomposition_gloo_world2(tmp_path):
    """geogrid shards by destination patch: two gloo ranks take the patches of a 2-task run, accumulate a synthetic
    categorical dataset onto their own patch with numpy, and the gathered result equals the single-task result."""
    script = tmp_path / "p.py"
    script.write_text(textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import numpy as np, torch, torch.distributed as dist
        from wps_b200.shard import patch_bounds
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        idim, jdim, ncat = 37, 29, 5
        rng = np.random.default_rng(3)
        # source pixels with a destination cell and a category each (what where_maps_to + the tile give)
        px = rng.integers(1, idim + 1, 20000); py = rng.integers(1, jdim + 1, 20000); cat = rng.integers(0, ncat, 20000)
        def accumulate(x0, x1, y0, y1):
            f = np.zeros((ncat, jdim, idim), np.float32)
            m = (px >= x0) & (px <= x1) & (py >= y0) & (py <= y1)
            np.add.at(f, (cat[m], py[m] - 1, px[m] - 1), 1.0)
            return f
        x0, x1, y0, y1 = patch_bounds(idim, jdim, world, rank)
        mine = torch.from_numpy(accumulate(x0, x1, y0, y1))
        dist.all_reduce(mine)     # stands for gather_whole_field_r: patches are disjoint, so the sum is the gather
        if rank == 0:
            assert np.array_equal(mine.numpy(), accumulate(1, idim, 1, jdim))
            print("OK", x0, x1, y0, y1)
        dist.destroy_process_group()
    """ % ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29613", str(script)], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK 1 37 1 14" in out.stdout


def test_fill_lev_table_options():
    """fill_lev / level_template / derived / flag_in_output (interp_option_module.F:271-311, 428-475, 736-812)."""
    from wps_b200.interp_option import get_constant_fill_lev, get_fill_src_level, read_interp_table
    t = read_interp_table("""
    ========================================
    name=TT ; fill_lev=200100:TT(100000) ; fill_lev=all:const(5.5)
    ========================================
    name=PRESSURE ; derived=yes ; z_dim_name=num_metgrid_levels ; level_template=TT
            fill_lev=all:PRES
            fill_lev=all:vertical_index; flag_in_output=FLAG_P
    ========================================
    """)
    assert t[0].fill_lev == [("200100", "TT(100000)"), ("all", "const(5.5)")] and not t[0].is_derived_field
    assert t[1].fill_lev == [("all", "PRES"), ("all", "vertical_index")]
    assert t[1].is_derived_field and t[1].level_template == "TT" and t[1].flag_in_output == "FLAG_P"
    assert get_constant_fill_lev("const(5.5)") == (0, 5.5) and get_constant_fill_lev("PRES")[0] == 1
    assert get_fill_src_level("TT(100000)") == ("TT", 100000) and get_fill_src_level("SOILHGT") == ("SOILHGT", 1)


@pytest.mark.parametrize("opt,npass", [(1, 1), (1, 3), (2, 1), (3, 2)])
def test_patchwise_smoothing_with_apron_equals_whole_domain(opt, npass):
    """geogrid shards by destination patch; smoothing a patch extended by smoother_apron() cells gives, on
    the patch, bit for bit what smoothing the whole domain gives (checked with the CPU oracle's smoothers; the GPU twin
    is tests/test_gpu_geogrid.py::test_patchwise_smoothing_on_device)."""
    from oracle import oracle as O
    from wps_b200.shard import apron_window, patch_bounds, smoother_apron
    idim, jdim, world = 61, 47, 4
    a = np.random.default_rng(3).uniform(-50, 3000, (jdim, idim)).astype(np.float32)
    whole = O.smooth(a, npass, opt)
    ap = smoother_apron(opt, npass)
    out = np.full_like(a, np.nan)
    for rank in range(world):
        x0, x1, y0, y1 = patch_bounds(idim, jdim, world, rank)
        ax0, ax1, ay0, ay1 = apron_window(x0, x1, y0, y1, idim, jdim, ap)
        s = O.smooth(a[ay0 - 1:ay1, ax0 - 1:ax1], npass, opt)
        out[y0 - 1:y1, x0 - 1:x1] = s[y0 - ay0:y1 - ay0 + 1, x0 - ax0:x1 - ax0 + 1]
    assert np.array_equal(out.view(np.int32), whole.view(np.int32))
    # one cell less is not enough (the test would be vacuous otherwise)
    if ap > 1:
        x0, x1, y0, y1 = patch_bounds(idim, jdim, world, 0)
        ax0, ax1, ay0, ay1 = apron_window(x0, x1, y0, y1, idim, jdim, ap - 1)
        s = O.smooth(a[ay0 - 1:ay1, ax0 - 1:ax1], npass, opt)
        assert not np.array_equal(s[y0 - ay0:y1 - ay0 + 1, x0 - ax0:x1 - ax0 + 1], whole[y0 - 1:y1, x0 - 1:x1])


def test_output_fields_order_follows_get_next_output_field():
    """storage_module.F:478-552 + storage_put_field :88-136, by hand: head nodes of new field names go to the FRONT of the
    list, levels stay sorted by descending vertical_level (secondary_cmp of time-dependent fields), mask fields
    (output_this_field = no) are skipped, and the levels of a field come out stacked in one array."""
    from wps_b200.metgrid import FgInput, output_fields
    from wps_b200 import capi

    def fg(name, lvl, val, mask=False):
        return FgInput(name, lvl, capi.M, np.full((2, 3), val, np.float32), np.zeros((2, 1), np.int32), None, False, mask)
    storage = {}
    for name, lvl, val, mask in [("TT", 85000, 1.0, False), ("LANDSEA", 200100, 9.0, True), ("TT", 200100, 2.0, False),
                                 ("UU", 100000, 3.0, False), ("TT", 100000, 4.0, False), ("UU", 50000, 5.0, False), ("PSFC", 200100, 6.0, False)]:
        storage[(name, lvl)] = fg(name, lvl, val, mask)
    out = list(output_fields(storage))
    assert [n for n, _ in out] == ["PSFC", "UU", "TT"]          # reverse of first storage, LANDSEA (mask field) skipped
    tt = out[2][1]
    assert tt.shape == (3, 2, 3) and [float(tt[k, 0, 0]) for k in range(3)] == [2.0, 4.0, 1.0]     # 200100, 100000, 85000
    assert [float(out[1][1][k, 0, 0]) for k in range(2)] == [3.0, 5.0]
