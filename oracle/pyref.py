"""Second, independent transliteration of WPS interp_module.F (v4.6.0) in pure Python -- TEST INFRASTRUCTURE.

Written separately from the C oracle (oracle/orc_interp.c), keeping the Fortran's recursive structure
(each operator tail-calls interp_sequence(idx+1) itself) and its linked-list queue, so that transcription slips
in one of the two restatements show up as disagreements (tests/test_oracle_crosscheck.py).  Every arithmetic
operation is done on numpy.float32 scalars to reproduce default-REAL arithmetic.  Small cases only (it is slow).
"""
from collections import deque

import numpy as np

F = np.float32
SIXTEEN_POINT, FOUR_POINT, N_NEIGHBOR, AVERAGE4, AVERAGE16, W_AVERAGE4, W_AVERAGE16, SEARCH = range(1, 9)


def sq(v):
    return v * v


def nint(x):
    x = float(x)
    return int(np.floor(x + 0.5)) if x >= 0 else -int(np.floor(-x + 0.5))


def floor(x):
    return int(np.floor(float(x)))


def ceiling(x):
    return int(np.ceil(float(x)))


def trunc(x):
    return int(float(x))


class Ctx:
    """array(start_x:end_x, start_y:end_y) [+ mask_array, maskval, mask_relational] as passed down the Fortran calls."""

    def __init__(self, array, sx, sy, msgval, interp_list, interp_opts, mask=None, maskval=None, rel=None):
        self.a = np.asarray(array, np.float32)
        self.sx, self.sy = sx, sy
        self.ey, self.ex = sy + self.a.shape[0] - 1, sx + self.a.shape[1] - 1
        self.msg = F(msgval)
        self.lst, self.opts = list(interp_list), list(interp_opts)
        self.mask = None if mask is None else np.asarray(mask, np.float32)
        self.maskval = None if maskval is None else F(maskval)
        self.rel = rel

    def A(self, i, j):
        return self.a[j - self.sy, i - self.sx]

    def M(self, i, j):
        return self.mask[j - self.sy, i - self.sx]

    def has_mask(self):
        return self.mask is not None and self.maskval is not None and self.rel is not None


def interp_sequence(c, xx, yy, idx):
    """interp_module.F:304-367 (idx is 1-based)."""
    m = c.lst[idx - 1]
    if m == 0:
        return c.msg
    if m == FOUR_POINT:
        return four_pt(c, xx, yy, idx + 1)
    if m == AVERAGE4:
        return four_pt_average(c, xx, yy, idx + 1)
    if m == W_AVERAGE4:
        return wt_four_pt_average(c, xx, yy, idx + 1)
    if m == N_NEIGHBOR:
        return nearest_neighbor(c, xx, yy, idx + 1)
    if m == SIXTEEN_POINT:
        return sixteen_pt(c, xx, yy, idx + 1)
    if m == SEARCH:
        return search_extrap(c, xx, yy, idx + 1)
    if m == AVERAGE16:
        return sixteen_pt_average(c, xx, yy, idx + 1)
    if m == W_AVERAGE16:
        return wt_sixteen_pt_average(c, xx, yy, idx + 1)
    raise ValueError(m)


def nearest_neighbor(c, xx, yy, idx):
    """:376-442"""
    ix, iy = nint(xx), nint(yy)
    if ix < c.sx or ix > c.ex:
        return c.msg
    if iy < c.sy or iy > c.ey:
        return c.msg
    if c.has_mask():
        if c.rel == '<' and c.M(ix, iy) < c.maskval:
            r = c.msg
        elif c.rel == '>' and c.M(ix, iy) > c.maskval:
            r = c.msg
        elif c.rel == ' ' and c.M(ix, iy) == c.maskval:
            r = c.msg
        else:
            r = c.A(ix, iy)
    else:
        r = c.A(ix, iy)
    if r == c.msg:
        r = interp_sequence(c, xx, yy, idx)
    return r


def search_extrap(c, xx, yy, idx):
    """:451-615 with a literal FIFO queue of (x, y, depth) records and a visited set."""
    if nint(xx) < c.sx or nint(xx) > c.ex or nint(yy) < c.sy or nint(yy) > c.ey:
        return c.msg
    visited = set()
    q = deque()
    maxdepth = c.opts[idx - 2]
    found_valid = False
    qd = [nint(xx), nint(yy), 0]
    q.append(tuple(qd))
    visited.add((qd[0], qd[1]))

    def valid(i, j):
        if c.has_mask():
            if c.rel == '>':
                return c.M(i, j) <= c.maskval and c.A(i, j) != c.msg
            if c.rel == '<':
                return c.M(i, j) >= c.maskval and c.A(i, j) != c.msg
            if c.rel == ' ':
                return c.M(i, j) != c.maskval and c.A(i, j) != c.msg
            return False
        return c.A(i, j) != c.msg

    i = j = 0
    while q and not found_valid:
        qd = list(q.popleft())
        i, j = qd[0], qd[1]
        if valid(i, j):
            found_valid = True
        for (ni, nj, inb) in ((i - 1, j, i - 1 >= c.sx), (i + 1, j, i + 1 <= c.ex), (i, j - 1, j - 1 >= c.sy), (i, j + 1, j + 1 <= c.ey)):
            if inb:
                if (ni, nj) not in visited:
                    if qd[2] < maxdepth:
                        qd[0], qd[1] = ni, nj
                        qd[2] = qd[2] + 1       # not reset between siblings
                        q.append(tuple(qd))
                        visited.add((ni, nj))
    if not found_valid:
        return c.msg
    distance = (F(i) - xx) * (F(i) - xx) + (F(j) - yy) * (F(j) - yy)
    r = c.A(i, j)
    while q:
        x, y, _ = q.popleft()
        if valid(x, y):
            d = (F(x) - xx) * (F(x) - xx) + (F(y) - yy) * (F(y) - yy)
            if d < distance:
                distance = d
                r = c.A(x, y)
    return r


def _unusable(c, i, j):
    if c.rel == '>':
        return c.M(i, j) > c.maskval
    if c.rel == '<':
        return c.M(i, j) < c.maskval
    return c.M(i, j) == c.maskval


def _avg4(c, xx, yy, idx, weighted):
    """:623-734 / :742-853"""
    ifx, icx, ify, icy = floor(xx), ceiling(xx), floor(yy), ceiling(yy)
    if weighted:
        fxfy = max(F(0.), F(1.0) - np.sqrt(sq(xx - F(ifx)) + sq(yy - F(ify)), dtype=np.float32))
        fxcy = max(F(0.), F(1.0) - np.sqrt(sq(xx - F(ifx)) + sq(yy - F(icy)), dtype=np.float32))
        cxfy = max(F(0.), F(1.0) - np.sqrt(sq(xx - F(icx)) + sq(yy - F(ify)), dtype=np.float32))
        cxcy = max(F(0.), F(1.0) - np.sqrt(sq(xx - F(icx)) + sq(yy - F(icy)), dtype=np.float32))
    else:
        fxfy = fxcy = cxfy = cxcy = F(1.0)
    if ifx < c.sx or icx > c.ex or ify < c.sy or icy > c.ey:
        if xx > F(c.sx) - F(0.5) and ifx < c.sx:
            ifx = icx = c.sx
        elif xx < F(c.ex) + F(0.5) and icx > c.ex:
            ifx = icx = c.ex
        ytest = (ifx < c.sy) if weighted else (ify < c.sy)      # :795 tests ifx in the weighted version
        if yy > F(c.sy) - F(0.5) and ytest:
            ify = icy = c.sy
        elif yy < F(c.ey) + F(0.5) and icy > c.ey:
            ify = icy = c.ey
        if ifx < c.sx or icx > c.ex or ify < c.sy or icy > c.ey:
            return c.msg
    pts = ((ifx, ify), (ifx, icy), (icx, ify), (icx, icy))
    w = [fxfy, fxcy, cxfy, cxcy]
    for k, (i, j) in enumerate(pts):
        if c.A(i, j) == c.msg or (c.has_mask() and _unusable(c, i, j)):
            w[k] = F(0.0)
    if all(v == 0.0 for v in w):
        return interp_sequence(c, xx, yy, idx)
    num = w[0] * c.A(*pts[0]) + w[1] * c.A(*pts[1]) + w[2] * c.A(*pts[2]) + w[3] * c.A(*pts[3])
    return num / (w[0] + w[1] + w[2] + w[3])


def four_pt_average(c, xx, yy, idx):
    return _avg4(c, xx, yy, idx, False)


def wt_four_pt_average(c, xx, yy, idx):
    return _avg4(c, xx, yy, idx, True)


def _avg16(c, xx, yy, idx, weighted):
    """:861-950 / :958-1047"""
    ifx, ify = floor(xx), floor(yy)
    if ifx < c.sx + 1 or ifx > c.ex - 2 or ify < c.sy + 1 or ify > c.ey - 2:
        return interp_sequence(c, xx, yy, idx)
    weights = {}
    sum_weight = F(0.0)
    for i in range(1, 5):
        for j in range(1, 5):
            px, py = ifx + 3 - i, ify + 3 - j
            if c.has_mask() and ((c.rel == '>' and c.M(px, py) > c.maskval) or (c.rel == '<' and c.M(px, py) < c.maskval) or
                                 (c.rel == ' ' and c.M(px, py) == c.maskval)):
                w = F(0.0)
            elif weighted:
                w = max(F(0.), F(2.0) - np.sqrt(sq(xx - F(px)) + sq(yy - F(py)), dtype=np.float32))
            else:
                w = F(1.0)
            if c.A(px, py) == c.msg:
                w = F(0.0)
            weights[(i, j)] = w
            sum_weight = sum_weight + w
    if sum_weight == 0.0:
        return interp_sequence(c, xx, yy, idx)
    s = F(0.0)
    for i in range(1, 5):
        for j in range(1, 5):
            s = s + weights[(i, j)] * c.A(ifx + 3 - i, ify + 3 - j)
    return s / sum_weight


def sixteen_pt_average(c, xx, yy, idx):
    return _avg16(c, xx, yy, idx, False)


def wt_sixteen_pt_average(c, xx, yy, idx):
    return _avg16(c, xx, yy, idx, True)


def four_pt(c, xx, yy, idx):
    """:1055-1171"""
    min_x, min_y, max_x, max_y = floor(xx), floor(yy), ceiling(xx), ceiling(yy)
    if min_x < c.sx or max_x > c.ex:
        return interp_sequence(c, xx, yy, idx)
    if min_y < c.sy or max_y > c.ey:
        return interp_sequence(c, xx, yy, idx)
    for (i, j) in ((min_x, min_y), (max_x, min_y), (min_x, max_y), (max_x, max_y)):
        if c.A(i, j) == c.msg or (c.has_mask() and _unusable(c, i, j)):
            return interp_sequence(c, xx, yy, idx)
    if min_x == max_x:
        if min_y == max_y:
            return c.A(min_x, min_y)
        return c.A(min_x, min_y) * (F(max_y) - yy) + c.A(min_x, max_y) * (yy - F(min_y))
    if min_y == max_y:
        return c.A(min_x, min_y) * (F(max_x) - xx) + c.A(max_x, min_y) * (xx - F(min_x))
    return (yy - F(min_y)) * (c.A(min_x, max_y) * (F(max_x) - xx) + c.A(max_x, max_y) * (xx - F(min_x))) + \
           (F(max_y) - yy) * (c.A(min_x, min_y) * (F(max_x) - xx) + c.A(max_x, min_y) * (xx - F(min_x)))


def oned(x, a, b, c, d):
    """:1342-1374"""
    r = F(0.)
    one, half = F(1.0), F(0.5)
    if x == 0.:
        r = b
    elif x == 1.:
        r = c
    if b * c != 0.:
        if a * d == 0.:
            if a == 0 and d == 0:
                r = b * (one - x) + c * x
            elif a != 0.:
                r = b + x * (half * (c - a) + x * (half * (c + a) - b))
            elif d != 0.:
                r = c + (one - x) * (half * (b - d) + (one - x) * (half * (b + d) - c))
        else:
            r = (one - x) * (b + x * (half * (c - a) + x * (half * (c + a) - b))) + \
                x * (c + (one - x) * (half * (b - d) + (one - x) * (half * (b + d) - c)))
    return r


def sixteen_pt(c, xx, yy, idx):
    """:1179-1334"""
    if trunc(xx) < c.sx or trunc(xx) > c.ex or trunc(yy) < c.sy or trunc(yy) > c.ey:
        return interp_sequence(c, xx, yy, idx)
    i = trunc(xx + F(0.00001))
    j = trunc(yy + F(0.00001))
    x = xx - F(i)
    y = yy - F(j)
    if abs(x) > F(0.0001) or abs(y) > F(0.0001):
        stl = {}
        is_masked = False
        for k in range(1, 5):
            kk = min(max(i + k - 2, c.sx), c.ex)
            for l in range(1, 5):
                ll = min(max(j + l - 2, c.sy), c.ey)
                v = c.A(kk, ll)
                if c.has_mask() and _unusable_rel(c, kk, ll):
                    is_masked = True
                if v == 0. and c.msg != 0.:
                    v = F(1.E-20)
                stl[(k, l)] = v
        for k in range(1, 5):
            for l in range(1, 5):
                if stl[(k, l)] == c.msg or (c.has_mask() and is_masked):
                    return interp_sequence(c, xx, yy, idx)
        a = oned(x, stl[(1, 1)], stl[(2, 1)], stl[(3, 1)], stl[(4, 1)])
        b = oned(x, stl[(1, 2)], stl[(2, 2)], stl[(3, 2)], stl[(4, 2)])
        cc = oned(x, stl[(1, 3)], stl[(2, 3)], stl[(3, 3)], stl[(4, 3)])
        d = oned(x, stl[(1, 4)], stl[(2, 4)], stl[(3, 4)], stl[(4, 4)])
        r = oned(y, a, b, cc, d)
        if r == F(1.E-20):
            r = F(0.)
        return r
    inb = c.sx <= i <= c.ex and c.sy <= j <= c.ey
    if c.has_mask():
        ok = inb and ((c.rel == '<' and c.M(i, j) >= c.maskval) or (c.rel == '>' and c.M(i, j) <= c.maskval) or
                      (c.rel == ' ' and c.M(i, j) != c.maskval)) and c.A(i, j) != c.msg
    else:
        ok = inb and c.A(i, j) != c.msg
    if ok:
        return c.A(i, j)
    return interp_sequence(c, xx, yy, idx)


def _unusable_rel(c, i, j):
    """the three explicit relational tests of sixteen_pt (:1245-1251)"""
    return (c.rel == '>' and c.M(i, j) > c.maskval) or (c.rel == '<' and c.M(i, j) < c.maskval) or \
           (c.rel == ' ' and c.M(i, j) == c.maskval)


def run(xx, yy, array, sx, sy, msgval, codes, opts, mask=None, maskval=None, rel=None):
    c = Ctx(array, sx, sy, msgval, list(codes) + [0], list(opts) + [0], mask, maskval, rel)
    return np.array([interp_sequence(c, F(x), F(y), 1) for x, y in zip(xx, yy)], np.float32)
