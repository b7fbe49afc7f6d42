"""CPU: the C oracle against the independent pure-Python transliteration (oracle/pyref.py) on seeded inputs,
for every operator, mask relational and edge situation.  Any disagreement means one of the two restatements of
interp_module.F is wrong."""
import numpy as np
import pytest

from oracle import pyref

CHAINS = [[1], [2], [3], [4], [5], [6], [7], [8], [1, 2, 4], [1, 2, 6, 7, 8], [2, 4, 8], [5, 3], [7, 2], [3, 8], [6, 8]]


def make_case(seed, nx=13, ny=11, frac_missing=0.0, zeros=0.0, msg=-1e30):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-5, 5, (ny, nx)).astype(np.float32)
    if zeros:
        a[rng.random(a.shape) < zeros] = 0.0
    if frac_missing:
        a[rng.random(a.shape) < frac_missing] = np.float32(msg)
    # points: inside, near nodes, near/outside every edge
    xx = np.concatenate([rng.uniform(-1.5, nx + 2.5, 120), np.arange(1, 8) + 0.00005, np.arange(1, 8).astype(float),
                         [0.49, 0.51, nx + 0.49, nx + 0.51, 1.0, float(nx)]]).astype(np.float32)
    yy = np.concatenate([rng.uniform(-1.5, ny + 2.5, 120), np.arange(1, 8) + 0.00002, np.arange(2, 9).astype(float),
                         [3.3, 0.6, 0.4, ny + 0.3, 1.0, float(ny)]]).astype(np.float32)
    return a, xx, yy


@pytest.mark.parametrize("chain", CHAINS)
@pytest.mark.parametrize("frac_missing,zeros,msg", [(0.0, 0.0, 1e20), (0.25, 0.05, -1e30), (0.6, 0.0, -1e30), (0.0, 0.2, 0.0)])
def test_operators_no_mask(orc, chain, frac_missing, zeros, msg):
    a, xx, yy = make_case(len(chain) * 7 + int(frac_missing * 10), frac_missing=frac_missing, zeros=zeros, msg=msg)
    opts = [4 if c == 8 else 0 for c in chain]
    for sx, sy in ((1, 1), (-2, 1), (5, -3)):
        got = orc.interp_sequence(xx + (sx - 1), yy + (sy - 1), a, sx, sy, np.float32(msg), (chain, opts))
        ref = pyref.run(xx + np.float32(sx - 1), yy + np.float32(sy - 1), a, sx, sy, msg, chain, opts)
        assert np.array_equal(got.view(np.int32), ref.view(np.int32)), (chain, sx, sy, np.flatnonzero(got != ref)[:5])


@pytest.mark.parametrize("chain", CHAINS)
@pytest.mark.parametrize("rel,maskval", [(" ", 1.0), ("<", 0.5), (">", 0.5)])
def test_operators_with_mask(orc, chain, rel, maskval):
    a, xx, yy = make_case(99 + len(chain), frac_missing=0.1, zeros=0.02)
    rng = np.random.default_rng(5)
    mask = (rng.random(a.shape) < 0.4).astype(np.float32)
    opts = [1200 if c == 8 else 0 for c in chain]
    got = orc.interp_sequence(xx, yy, a, 1, 1, np.float32(-1e30), (chain, opts), mask=mask, rel=rel, maskval=maskval)
    ref = pyref.run(xx, yy, a, 1, 1, -1e30, chain, opts, mask=mask, maskval=maskval, rel=rel)
    assert np.array_equal(got.view(np.int32), ref.view(np.int32)), (chain, rel)


@pytest.mark.parametrize("depth", [1, 2, 3, 5, 8, 1200])
def test_search_depth_quirk(orc, depth):
    """the cumulative depth counter (interp_module.F:521-561) in both restatements"""
    rng = np.random.default_rng(depth)
    a = np.full((15, 17), -1e30, np.float32)
    a[rng.random(a.shape) < 0.06] = 1.0
    a *= np.arange(a.size, dtype=np.float32).reshape(a.shape) + 1
    xx = rng.uniform(0.6, 17.4, 150).astype(np.float32)
    yy = rng.uniform(0.6, 15.4, 150).astype(np.float32)
    got = orc.interp_sequence(xx, yy, a, 1, 1, np.float32(-1e30), ([8], [depth]))
    ref = pyref.run(xx, yy, a, 1, 1, -1e30, [8], [depth])
    assert np.array_equal(got.view(np.int32), ref.view(np.int32))
