"""Pins the tile decode against what the reference itself holds: geogrid/src/read_geogrid.c, the one C file on the path.

 * not-gpu: the committed golden vectors (tests/golden/read_geogrid.npz, made by running the reference's own code) equal
   oracle/_ref/libread_geogrid.so when that library is present, and the numpy statement of the decode rule used by the other
   tests reproduces them -- including the reference's sign rule `ival > 2^(8w-1)` (the most negative word stays positive)
   and its 4-byte behaviour (`1 << 32`: no sign correction at all).
 * gpu: wpsgpu_geogrid_decode_tile (k_geo_decode, the kernel behind wpsgpu_geogrid_add_tile_raw) reproduces the golden
   vectors bit for bit for wordsize 1-4 x signed x endian, with and without a scale factor, and the TOP_BOTTOM flip."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "read_geogrid.npz"))
CASES = [(w, s, e) for w in (1, 2, 3, 4) for s in (0, 1) for e in (0, 1)]


def numpy_decode(raw, wordsize, signed, endian):
    """read_geogrid.c:91-128 restated: words assembled MSB first, sign correction only when ival > 2^(8w-1), w < 4"""
    b = raw.astype(np.int64)
    if endian == 1:
        b = b[:, ::-1]
    val = np.zeros(b.shape[0], np.int64)
    for k in range(wordsize):
        val = (val << 8) | b[:, k]
    if wordsize == 4:
        val = np.where(val >= (1 << 31), val - (1 << 32), val)       # (int) of the assembled word; `ival -= (1 << 32)` subtracts 0
    elif signed:
        val = np.where(val > (1 << (8 * wordsize - 1)), val - (1 << (8 * wordsize)), val)
    return val.astype(np.float32)


@pytest.mark.parametrize("wordsize,signed,endian", CASES)
def test_golden_vectors_match_numpy_rule(wordsize, signed, endian):
    key = "w%d_s%d_e%d" % (wordsize, signed, endian)
    dec = numpy_decode(GOLD[key + "_raw"], wordsize, signed, endian)
    assert dec.tobytes() == GOLD[key + "_dec"].tobytes(), key


def test_golden_vectors_match_reference_library():
    path = os.path.join(ROOT, "oracle", "_ref", "libread_geogrid.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference; `make -C oracle ref`)")
    from tests.golden.make_golden_read_geogrid import ref_decode
    lib = C.CDLL(path)
    for wordsize, signed, endian in CASES:
        key = "w%d_s%d_e%d" % (wordsize, signed, endian)
        raw = GOLD[key + "_raw"]
        assert ref_decode(lib, raw, raw.shape[0], wordsize, signed, endian).tobytes() == GOLD[key + "_dec"].tobytes(), key
    raw = GOLD["w2_s1_e0_raw"]
    assert ref_decode(lib, raw, raw.shape[0], 2, 1, 0, scale=0.1).tobytes() == GOLD["w2_s1_e0_scaled_dec"].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("wordsize,signed,endian", CASES)
def test_device_decode_matches_reference(handle, wordsize, signed, endian):
    key = "w%d_s%d_e%d" % (wordsize, signed, endian)
    raw, dec = GOLD[key + "_raw"], GOLD[key + "_dec"]
    n = raw.shape[0]
    nx, ny = 4, n // 4
    raw = np.ascontiguousarray(raw[:nx * ny])
    ref = dec[:nx * ny].reshape(1, ny, nx)
    out = handle.geogrid_decode_tile(raw, (1, ny, nx), wordsize, signed, endian)
    assert out.tobytes() == ref.tobytes(), "device decode differs from read_geogrid.c for " + key
    # row_order = top_bottom: the reference flips rows after read_geogrid (source_data_module.F:2887-2898)
    out = handle.geogrid_decode_tile(raw, (1, ny, nx), wordsize, signed, endian, row_order=2)
    assert out.tobytes() == np.ascontiguousarray(ref[:, ::-1, :]).tobytes(), "row flip, " + key


@pytest.mark.gpu
def test_device_decode_scale_factor(handle):
    raw, dec = GOLD["w2_s1_e0_raw"], GOLD["w2_s1_e0_scaled_dec"]
    n = raw.shape[0]
    out = handle.geogrid_decode_tile(np.ascontiguousarray(raw), (1, 1, n), 2, 1, 0, scalefactor=0.1)
    assert out.reshape(-1).tobytes() == dec.tobytes()
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libread_geogrid.so")):
        # the library itself, on the box: fresh random bytes, not only the committed vectors
        from tests.golden.make_golden_read_geogrid import ref_decode
        lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libread_geogrid.so"))
        rng = np.random.default_rng(5)
        for wordsize, signed, endian in CASES:
            raw = rng.integers(0, 256, (3000, wordsize), dtype=np.uint8)
            ref = ref_decode(lib, raw, 3000, wordsize, signed, endian)
            out = handle.geogrid_decode_tile(raw, (1, 30, 100), wordsize, signed, endian)
            assert out.reshape(-1).tobytes() == ref.tobytes(), (wordsize, signed, endian)
