"""Generates tests/golden/read_geogrid.npz with the REFERENCE's own code: oracle/_ref/libread_geogrid.so is
/root/reference/geogrid/src/read_geogrid.c compiled as it lies (recipe: `make -C oracle ref`).  For every
wordsize 1..4 x signed x endian it decodes a fixed byte pattern -- random bytes plus every sign-boundary word -- through
read_geogrid() and stores bytes and decoded floats.  Run in the build container (needs /root/reference):
    python tests/golden/make_golden_read_geogrid.py"""
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def ref_decode(lib, raw, n, wordsize, signed, endian, scale=1.0):
    with tempfile.NamedTemporaryFile(delete=False) as f:
        f.write(raw.tobytes())
        name = f.name
    out = np.zeros(n, np.float32)
    fn = name.encode()
    i = lambda v: C.byref(C.c_int(v))
    status = C.c_int(0)
    lib.read_geogrid(C.c_char_p(fn), i(len(fn)), out.ctypes.data_as(C.c_void_p), i(n), i(1), i(1), i(signed), i(endian),
                     C.byref(C.c_float(scale)), i(wordsize), C.byref(status))
    os.unlink(name)
    assert status.value == 0
    return out


def patterns(wordsize, rng, n_random=4000):
    """random words + the words around every sign boundary, as bytes in BIG-endian order [n][wordsize]"""
    top = 1 << (8 * wordsize)
    special = [0, 1, 2, (top >> 1) - 1, top >> 1, (top >> 1) + 1, top - 2, top - 1, 0x7f, 0x80, 0x81, 0xff]
    special = [v % top for v in special]
    vals = np.concatenate([np.array(special, np.uint64), rng.integers(0, top, n_random, dtype=np.uint64)])
    return np.stack([((vals >> np.uint64(8 * b)) & np.uint64(0xff)).astype(np.uint8) for b in range(wordsize - 1, -1, -1)], axis=-1)


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libread_geogrid.so"))
    rng = np.random.default_rng(20261020)
    out = {}
    for wordsize in (1, 2, 3, 4):
        be = patterns(wordsize, rng)
        for endian in (0, 1):
            raw = be if endian == 0 else be[:, ::-1].copy()
            for signed in (0, 1):
                key = "w%d_s%d_e%d" % (wordsize, signed, endian)
                out[key + "_raw"] = raw
                out[key + "_dec"] = ref_decode(lib, raw, raw.shape[0], wordsize, signed, endian)
    raw = out["w2_s1_e0_raw"]
    out["w2_s1_e0_scaled_dec"] = ref_decode(lib, raw, raw.shape[0], 2, 1, 0, scale=0.1)
    np.savez_compressed(os.path.join(HERE, "read_geogrid.npz"), **out)
    print("wrote read_geogrid.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
