"""GPU parity: fill_missing_levels / fill_field / create_level (process_domain_module.F:2748-3332), the first of the
SURVEY 8(f) rows -- bit-exact against the oracle's literal restatement of the reference loops."""
import numpy as np
import pytest

from tests.util import assert_bitexact, bool_to_bits, bits_to_bool

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nx,ny,nlev,p_valid", [(70, 33, 12, 0.6), (129, 17, 35, 0.3), (31, 5, 2, 0.5), (64, 8, 9, 0.95), (40, 9, 50, 0.4), (97, 6, 64, 0.2),
                                                  (33, 7, 70, 0.5)])
def test_fill_missing_levels_vs_oracle(orc, handle, nx, ny, nlev, p_valid):
    from wps_b200 import capi
    rng = np.random.default_rng(nx * 1000 + nlev)
    levels = sorted(rng.choice(np.arange(1000, 110000, 250), nlev, replace=False).tolist())
    if nlev > 4:
        levels[-1] = 200100
    r = rng.uniform(200, 320, (nlev, ny, nx)).astype(np.float32)
    valid_b = rng.random((nlev, ny, nx)) < p_valid
    valid_b[0, :, : nx // 3] = False          # columns without a valid bottom level
    valid_b[-1, :, nx // 2:] = False          # columns without a valid top level
    valid_b[:, 0, 0] = False                  # a column with nothing at all
    valid = np.stack([bool_to_bits(valid_b[k]) for k in range(nlev)])
    r_ref, v_ref = orc.fill_missing_levels(levels, r, valid)
    r_dev = [r[k].copy() for k in range(nlev)]
    v_dev = [valid[k].copy() for k in range(nlev)]
    handle.fill_missing_levels(levels, r_dev, v_dev)
    for k in range(nlev):
        assert (v_dev[k] == v_ref[k]).all(), "valid_mask of level %d" % k
        assert_bitexact(r_dev[k], r_ref[k], "level %d" % k)
    filled = np.stack([bits_to_bool(v_dev[k], nx) for k in range(nlev)]) & ~valid_b
    assert filled.any() == (nlev > 2)     # with two levels no point has both a lower and an upper neighbour


def test_fill_missing_levels_idempotent_and_single_level(handle):
    """A second pass finds nothing left to fill between valid levels; a 2-d field (one level) is left alone."""
    rng = np.random.default_rng(4)
    nx, ny, nlev = 50, 20, 8
    levels = [100 * (k + 1) for k in range(nlev)]
    r = [rng.uniform(0, 1, (ny, nx)).astype(np.float32) for _ in range(nlev)]
    v = [bool_to_bits(rng.random((ny, nx)) < 0.5) for _ in range(nlev)]
    handle.fill_missing_levels(levels, r, v)
    r1, v1 = [a.copy() for a in r], [a.copy() for a in v]
    handle.fill_missing_levels(levels, r, v)
    for k in range(nlev):
        assert (v[k] == v1[k]).all()
        assert_bitexact(r[k], r1[k], "second pass, level %d" % k)
    one_r, one_v = [r[0].copy()], [np.zeros_like(v[0])]
    handle.fill_missing_levels([500], one_r, one_v)
    assert (one_v[0] == 0).all()


def test_create_level_const_and_copy(handle):
    from wps_b200 import capi
    rng = np.random.default_rng(5)
    for nx, ny in ((70, 9), (64, 4), (33, 3)):
        r = np.full((ny, nx), -1.0, np.float32)
        v = np.zeros((ny, capi.nwords(nx)), np.int32)
        handle.create_level(r, v, fill_const=2.5)
        assert (r == 2.5).all()
        assert bits_to_bool(v, nx).all()
        tail = bits_to_bool(v, capi.nwords(nx) * 32)[:, nx:]
        assert not tail.any()                     # bitarray_set is called for i = 1..nx only
        src = rng.uniform(0, 1, (ny, nx)).astype(np.float32)
        sv = bool_to_bits(rng.random((ny, nx)) < 0.5)
        r2 = np.zeros_like(r)
        v2 = np.zeros_like(v)
        handle.create_level(r2, v2, src=src, src_valid=sv)
        assert_bitexact(r2, src, "copied level")
        assert (v2 == sv).all()


def test_fill_field_rules_through_the_host_mirror(orc, handle):
    """fill_missing_levels end to end on a small storage: TT lacks two levels that RH has (created by
    fill_lev=all:const, then nothing to interpolate there), GHT has holes that the vertical interpolation fills,
    PSFC-like 2-d fields are untouched, and fill_lev=200100:PSFC(200100) copies a surface level with its valid bits."""
    from wps_b200 import capi
    from wps_b200.interp_option import read_interp_table
    from wps_b200.metgrid import FgInput, MetgridDomain
    nx, ny = 40, 30
    rng = np.random.default_rng(6)
    lat = np.tile(np.linspace(30, 40, ny, dtype=np.float32)[:, None], (1, nx))
    lon = np.tile(np.linspace(-100, -90, nx, dtype=np.float32)[None, :], (ny, 1))
    latu, lonu = np.zeros((ny, nx + 1), np.float32), np.zeros((ny, nx + 1), np.float32)
    latv, lonv = np.zeros((ny + 1, nx), np.float32), np.zeros((ny + 1, nx), np.float32)
    dom = MetgridDomain(handle, lat, lon, latu, lonu, latv, lonv, None)
    table = read_interp_table("""
    ========================================
    name=TT ; fill_lev=200100:PSFC(200100) ; fill_lev=all:const(-7.)
    ========================================
    name=GHT
    ========================================
    name=PSFC
    ========================================
    """)
    nw = capi.nwords(nx)
    full = bool_to_bits(np.ones((ny, nx), bool))

    def fld(name, lev, valid=None):
        return FgInput(name, lev, capi.M, rng.uniform(200, 300, (ny, nx)).astype(np.float32),
                       full.copy() if valid is None else valid, np.zeros((ny, nw), np.int32))

    storage = {}
    for lev in (100000, 85000, 70000, 50000):
        storage[("RH", lev)] = fld("RH", lev)
    for lev in (100000, 70000):
        storage[("TT", lev)] = fld("TT", lev)
    holes = bool_to_bits(rng.random((ny, nx)) < 0.5)
    storage[("GHT", 100000)] = fld("GHT", 100000)
    storage[("GHT", 85000)] = fld("GHT", 85000, holes.copy())
    storage[("GHT", 70000)] = fld("GHT", 70000)
    psfc_valid = bool_to_bits(rng.random((ny, nx)) < 0.8)
    storage[("PSFC", 200100)] = fld("PSFC", 200100, psfc_valid.copy())
    ght_before = {l: (storage[("GHT", l)].r_arr.copy(), storage[("GHT", l)].valid_mask.copy()) for l in (100000, 85000, 70000)}
    dom.fill_missing_levels(storage, table)
    # TT: surface level copied from PSFC with its valid bits, every other level of the union filled with the constant
    assert_bitexact(storage[("TT", 200100)].r_arr, storage[("PSFC", 200100)].r_arr, "TT(200100) <- PSFC")
    assert (storage[("TT", 200100)].valid_mask == psfc_valid).all()
    for lev in (85000, 50000):
        assert (storage[("TT", lev)].r_arr == np.float32(-7.0)).all() and (storage[("TT", lev)].valid_mask == full).all()
    # GHT: holes at 85000 interpolated between 70000 and 100000 exactly like the oracle does
    lv = [70000, 85000, 100000]
    r_ref, v_ref = orc.fill_missing_levels(lv, np.stack([ght_before[l][0] for l in lv]), np.stack([ght_before[l][1] for l in lv]))
    for k, l in enumerate(lv):
        assert_bitexact(storage[("GHT", l)].r_arr, r_ref[k], "GHT %d" % l)
        assert (storage[("GHT", l)].valid_mask == v_ref[k]).all()
    assert (storage[("GHT", 85000)].valid_mask == full).all()
    # GHT has no table rule creating levels, RH has no entry at all: their level sets are unchanged
    assert sorted(l for (n, l) in storage if n == "GHT") == lv
    assert sorted(l for (n, l) in storage if n == "RH") == [50000, 70000, 85000, 100000]
