"""GPU: the host-side mirror of metgrid's record loop (wps_b200/metgrid.py) on a synthetic GFS-shaped FILE written in
the intermediate format, against a per-slab oracle loop that follows process_intermediate_fields by hand."""
import numpy as np
import pytest

from tests.test_host_logic import TBL
from tests.util import assert_bitexact, bits_to_bool, conus_like, landsea, synth_field

pytestmark = pytest.mark.gpu


def make_file(tmp_path, name, fields, seed0, nx=360, ny=181):
    from wps_b200.intermediate import MetData, read_records, write_records
    ls = landsea(nx, ny)
    recs = []
    for k, (f, lvl, kind) in enumerate(fields):
        if kind == "landsea":
            slab = ls.copy()
        elif kind == "land":
            slab = synth_field(nx, ny, seed0 + k)
            slab[ls == 0] = -1e30
        else:
            slab = synth_field(nx, ny, seed0 + k, kind=kind)
        recs.append(MetData(f, lvl, slab, startlat=90.0, startlon=0.0, deltalat=-1.0, deltalon=1.0))
    p = str(tmp_path / name)
    write_records(p, recs)
    return read_records(p)


def test_record_loop_matches_oracle(orc, handle, tmp_path):
    from wps_b200 import capi
    from wps_b200.interp_option import find_entry, read_interp_table
    from wps_b200.metgrid import MetgridDomain, halo_slab
    orc.set_transc_mode(1)
    nx, ny = 64, 52
    odom, gdom = conus_like(nx, ny, 40000.0, 30.0, 60.0, -98.0, 36.0, -92.0)
    g = {1: orc.get_lat_lon_fields(odom, 1, 1, nx, ny, orc.M), 2: orc.get_lat_lon_fields(odom, 1, 1, nx + 1, ny, orc.U),
         3: orc.get_lat_lon_fields(odom, 1, 1, nx, ny + 1, orc.V)}
    rng = np.random.default_rng(1)
    landmask = (rng.random((ny, nx)) < 0.6).astype(np.float32)
    table = read_interp_table(TBL)
    fields = [("LANDSEA", 200100.0, "landsea"), ("T", 85000.0, "smooth"), ("T", 50000.0, "smooth"), ("UU", 85000.0, "wind"),
              ("VV", 85000.0, "wind"), ("ST000010", 200100.0, "land"), ("SKINTEMP", 200100.0, "smooth"), ("PSFC", 200100.0, "smooth")]
    recs = make_file(tmp_path, "FILE:2026-10-19_00", fields, 10)
    dom = MetgridDomain(handle, g[1][0], g[1][1], g[2][0], g[2][1], g[3][0], g[3][1], landmask)
    storage = dom.process_intermediate_fields(recs, table, "FILE")
    assert ("TT", 85000) in storage and ("T", 85000) not in storage          # output_name rename
    # ---- the same loop by hand through the oracle
    src = orc.push_source_projection(orc.PROJ_LATLON, dlat=-1.0, dlon=1.0, known_x=1.0, known_y=1.0, known_lat=90.0, known_lon=0.0,
                                     earth_radius=np.float32(6371.229) * np.float32(1000.0))
    masks = {"LANDSEA.mask": halo_slab(recs[0])[0]}
    for fg in recs:
        idx = find_entry(table, fg.field, "FILE")
        name = fg.field
        if table[idx].output_name != " ":
            name = table[idx].output_name
            idx = find_entry(table, name, "FILE")
        e = table[idx]
        stag = e.output_stagger if e.output_stagger in (2, 3) else 1
        fo = orc.field_opts(e.interp_method, np.float32(e.missing_value), np.float32(e.fill_missing), e.masked,
                            (e.interp_mask_val, e.interp_land_mask_val, e.interp_water_mask_val),
                            e.interp_mask_relational + e.interp_land_mask_relational + e.interp_water_mask_relational)
        m = [masks.get(nm + ".mask") if nm != " " else None for nm in (e.interp_mask, e.interp_land_mask, e.interp_water_mask)]
        slab, minx, bdr = halo_slab(fg)
        r_ref, np_ref = orc.interp_met_field(src, fo, stag, g[stag][0], g[stag][1], 1, 1, (1, nx, 1, ny), slab, minx, 1, bdr,
                                             mask=m[0], land_mask=m[1], water_mask=m[2], landmask=landmask if stag == 1 else None)
        got = storage[(name, int(round(fg.xlvl)))]
        assert (got.modified_mask == np_ref).all(), name
        assert (got.valid_mask == np_ref).all(), name
        assert_bitexact(got.r_arr, r_ref, "record loop " + name)
    # ---- a second, higher-priority source that only partly covers the field: valid_mask carry-over (:1225-1237, :1365)
    recs2 = make_file(tmp_path, "SST:2026-10-19_00", [("T", 85000.0, "smooth")], 99)
    recs2[0].slab[:, :180] = 1e20
    before = storage[("TT", 85000)].r_arr.copy()
    dom.process_intermediate_fields(recs2, table, "SST", storage)
    fo = orc.field_opts("sixteen_pt+four_pt+average_4pt", np.float32(1e20), np.float32(0.0))
    slab, minx, bdr = halo_slab(recs2[0])
    valid0 = storage[("TT", 85000)].valid_mask    # already merged; reconstruct the pre-merge valid mask = all set for TT
    r_ref, np_ref = orc.interp_met_field(src, fo, 1, g[1][0], g[1][1], 1, 1, (1, nx, 1, ny), slab, minx, 1, bdr, landmask=landmask,
                                         r_arr=before.copy(), valid_mask=np.full_like(valid0, -1))
    assert_bitexact(storage[("TT", 85000)].r_arr, r_ref, "second source")
    assert (storage[("TT", 85000)].modified_mask == np_ref).all()
    # ---- met_to_map on the interpolated winds (process_domain_module.F:802-834)
    u0, v0 = storage[("UU", 85000)].r_arr.copy(), storage[("VV", 85000)].r_arr.copy()
    dom.rotate_winds(storage, gdom, -1)
    uo, vo = orc.metmap_xform(odom, u0, storage[("UU", 85000)].modified_mask, v0, storage[("VV", 85000)].modified_mask, g[2][1], g[3][1], -1)
    from tests.util import ulp_diff
    assert ulp_diff(storage[("UU", 85000)].r_arr, uo).max() <= 1 and ulp_diff(storage[("VV", 85000)].r_arr, vo).max() <= 1


def test_cpp_host_side_matches_python_mirror(orc, handle, tmp_path):
    """include/wpshost.h: the C++ restatement of the host side (METGRID.TBL parser, intermediate-file reader, record
    loop with storage / mask carry-over, fill_missing_levels) drives the same C ABI -- handing the records over in file
    byte order -- and must reproduce the Python mirror (itself checked against the oracle above) bit for bit."""
    from wps_b200 import hostlib
    from wps_b200.interp_option import read_interp_table
    from wps_b200.metgrid import MetgridDomain
    orc.set_transc_mode(1)
    nx, ny = 64, 52
    odom, gdom = conus_like(nx, ny, 40000.0, 30.0, 60.0, -98.0, 36.0, -92.0)
    g = {1: orc.get_lat_lon_fields(odom, 1, 1, nx, ny, orc.M), 2: orc.get_lat_lon_fields(odom, 1, 1, nx + 1, ny, orc.U),
         3: orc.get_lat_lon_fields(odom, 1, 1, nx, ny + 1, orc.V)}
    rng = np.random.default_rng(1)
    landmask = (rng.random((ny, nx)) < 0.6).astype(np.float32)
    tbl = TBL + """
    name=GHT ; interp_option=four_pt+average_4pt ; fill_lev=200100:PSFC(200100)
    ========================================
    name=PRES ; derived=yes ; fill_lev=all:vertical_index ; level_template=TT
    ========================================
    """
    table = read_interp_table(tbl)
    fields = [("LANDSEA", 200100.0, "landsea"), ("T", 85000.0, "smooth"), ("T", 50000.0, "smooth"), ("T", 30000.0, "smooth"),
              ("UU", 85000.0, "wind"), ("VV", 85000.0, "wind"), ("ST000010", 200100.0, "land"), ("SKINTEMP", 200100.0, "smooth"),
              ("PSFC", 200100.0, "smooth"), ("HGT", 85000.0, "smooth"), ("HGT", 30000.0, "smooth")]
    recs = make_file(tmp_path, "FILE:2026-10-19_00", fields, 10)
    recs2 = make_file(tmp_path, "SST:2026-10-19_00", [("T", 50000.0, "smooth")], 99)
    # make the second source partial: re-write its file with missing values over half the globe
    from wps_b200.intermediate import write_records
    recs2[0].slab[:, :180] = 1e20
    write_records(str(tmp_path / "SST:2026-10-19_00"), recs2)
    # ---- Python mirror
    dom = MetgridDomain(handle, g[1][0], g[1][1], g[2][0], g[2][1], g[3][0], g[3][1], landmask)
    storage = dom.process_intermediate_fields(recs, table, "FILE")
    dom.process_intermediate_fields(recs2, table, "SST", storage)
    dom.create_derived_fields(storage, table)
    dom.fill_missing_levels(storage, table)
    # ---- C++ host side on the same files
    hd = hostlib.HostDomain(0, g[1][0], g[1][1], g[2][0], g[2][1], g[3][0], g[3][1], landmask)
    try:
        assert hd.set_table(tbl) == len(table)
        assert hd.process_file(str(tmp_path / "FILE:2026-10-19_00"), "FILE") == len(fields)
        assert hd.process_file(str(tmp_path / "SST:2026-10-19_00"), "SST") == 1
        hd.create_derived_fields()
        hd.fill_missing_levels()
        got = {(n, l) for (n, l, _, _, _) in hd.fields()}
        assert got == set(storage.keys()), (sorted(got ^ set(storage.keys())))
        assert ("GHT", 200100) in got and ("PRES", 50000) in got and ("TT", 200100) not in got
        for (name, level), f in storage.items():
            r, v = hd.get_field(name, level)
            assert (v == f.valid_mask).all(), (name, level)
            assert_bitexact(r, f.r_arr, "C++ host side %s %d" % (name, level))
    finally:
        hd.close()


def test_two_records_of_one_field_level_in_a_file(orc, handle, tmp_path):
    """two records of a file that resolve to the same (field, level).  The reference processes records one
    after another and merges valid_mask after each (process_domain_module.F:1225-1365): the second interpolation sees the
    first one's points as valid.  Both host mirrors flush before queueing the duplicate; the result must equal the
    sequential oracle loop (first record partly missing, second record fills only what it has)."""
    from wps_b200.interp_option import read_interp_table
    from wps_b200.metgrid import MetgridDomain, halo_slab
    orc.set_transc_mode(1)
    nx, ny = 58, 47
    odom, gdom = conus_like(nx, ny, 40000.0, 30.0, 60.0, -98.0, 36.0, -92.0)
    g = {1: orc.get_lat_lon_fields(odom, 1, 1, nx, ny, orc.M), 2: orc.get_lat_lon_fields(odom, 1, 1, nx + 1, ny, orc.U),
         3: orc.get_lat_lon_fields(odom, 1, 1, nx, ny + 1, orc.V)}
    table = read_interp_table(TBL)
    recs = make_file(tmp_path, "FILE:2026-10-19_00", [("T", 85000.0, "smooth"), ("T", 85000.0, "wind"), ("T", 70000.0, "smooth")], 300)
    recs[0].slab[:, 200:300] = 1e20          # first record: a band of missing values over the domain's longitudes
    recs[1].slab[:90, :] = 1e20              # second record: missing elsewhere
    dom = MetgridDomain(handle, g[1][0], g[1][1], g[2][0], g[2][1], g[3][0], g[3][1], None)
    storage = dom.process_intermediate_fields(recs, table, "FILE")
    src = orc.push_source_projection(orc.PROJ_LATLON, dlat=-1.0, dlon=1.0, known_x=1.0, known_y=1.0, known_lat=90.0, known_lon=0.0,
                                     earth_radius=np.float32(6371.229) * np.float32(1000.0))
    fo = orc.field_opts("sixteen_pt+four_pt+average_4pt", np.float32(1e20), np.float32(0.0))
    slab, minx, bdr = halo_slab(recs[0])
    r1, n1 = orc.interp_met_field(src, fo, 1, g[1][0], g[1][1], 1, 1, (1, nx, 1, ny), slab, minx, 1, bdr)
    slab, minx, bdr = halo_slab(recs[1])
    r2, n2 = orc.interp_met_field(src, fo, 1, g[1][0], g[1][1], 1, 1, (1, nx, 1, ny), slab, minx, 1, bdr, r_arr=r1.copy(), valid_mask=n1)
    got = storage[("TT", 85000)]
    assert (got.modified_mask == n2).all()
    assert (got.valid_mask == (n1 | n2)).all()
    assert_bitexact(got.r_arr, r2, "second record of the same (field, level)")
    assert (n1 != n2).any() and (bits_to_bool(n1, nx) & ~bits_to_bool(n2, nx)).any()
