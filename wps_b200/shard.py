"""Multi-GPU sharding of the metgrid path: units are 2-D slabs (field x level x time); one process per GPU, no
data-path collective.  Two couplings are respected: UU and VV of the same level stay on the same
rank (wind rotation needs both), and the "<NAME>.mask" slabs of a file are needed wherever a masked field is."""
from typing import Dict, List, Sequence, Tuple


def shard_times(n_times: int, rank: int, world: int) -> List[int]:
    """Round-robin over FILE times (the unit the bench uses: one time per GPU per step)."""
    return list(range(rank, n_times, world))


def shard_slabs(keys: Sequence[Tuple[str, int]], world: int, pairs=(("UU", "VV"),)) -> Dict[int, List[int]]:
    """Assign slab indices to ranks by level so that the components named in `pairs` share a rank.
    keys[i] = (field, level).  Returns {rank: [slab indices]} with balanced counts."""
    pair_of = {}
    for a, b in pairs:
        pair_of[a] = a + "|" + b
        pair_of[b] = a + "|" + b
    groups: Dict[Tuple[str, int], List[int]] = {}
    for i, (f, l) in enumerate(keys):
        groups.setdefault((pair_of.get(f, f), l), []).append(i)
    out: Dict[int, List[int]] = {r: [] for r in range(world)}
    load = [0] * world
    for g in sorted(groups.values(), key=len, reverse=True):
        r = load.index(min(load))
        out[r].extend(g)
        load[r] += len(g)
    for r in out:
        out[r].sort()
    return out


# ---- geogrid: destination-patch decomposition (the reference's own MPI decomposition) ---------------------------
def proc_grid(nprocs: int) -> Tuple[int, int]:
    """(nproc_x, nproc_y) as square as possible, first minimum wins -- parallel_module.F:62-74."""
    mini, npx, npy = 2 * nprocs, 1, nprocs
    for m in range(1, nprocs + 1):
        if nprocs % m == 0:
            n = nprocs // m
            if abs(m - n) < mini:
                mini, npx, npy = abs(m - n), m, n
    return npx, npy


def _task_1d(p: int, ds: int, de: int, nproc: int) -> int:
    """One direction of task_for_point (parallel_module.F:218-267, RSL_LITE's rule): the remainder rows are split
    between the first rem/2 and the last rem - rem/2 tasks.  Fortran integer division == // for these non-negatives."""
    i, ids, ide = p - 1, ds - 1, de - 1
    idim = ide - ids + 1
    i = min(max(i, ids), ide)
    rem = idim % nproc
    a = (rem // 2) * ((idim // nproc) + 1)
    b = a + (nproc - rem) * (idim // nproc)
    if i - ids < a:
        return (i - ids) // ((idim // nproc) + 1)
    if i - ids < b:
        return (a // ((idim // nproc) + 1)) + (i - a - ids) // ((b - a) // (nproc - rem))
    return (a // ((idim // nproc) + 1)) + (b - a - ids) // ((b - a) // (nproc - rem)) + (i - b - ids) // ((idim // nproc) + 1)


def task_for_point(i: int, j: int, ids: int, ide: int, jds: int, jde: int, npx: int, npy: int) -> Tuple[int, int]:
    return _task_1d(i, ids, ide, npx), _task_1d(j, jds, jde, npy)


def patch_bounds(idim: int, jdim: int, nprocs: int, rank: int) -> Tuple[int, int, int, int]:
    """(my_minx, my_maxx, my_miny, my_maxy), 1-based inclusive -- parallel_get_tile_dims, parallel_module.F:103-148.
    rank -> (my_x, my_y) = (rank mod nproc_x, rank / nproc_x) (:78-79).  One rank = one GPU: every GPU accumulates the
    source tiles overlapping its own patch (+ the caller's halo), no exchange step."""
    npx, npy = proc_grid(nprocs)
    my_x, my_y = rank % npx, rank // npx
    xs = [i for i in range(1, idim + 1) if _task_1d(i, 1, idim, npx) == my_x]
    ys = [j for j in range(1, jdim + 1) if _task_1d(j, 1, jdim, npy) == my_y]
    if not xs or not ys:
        return (-1, -1, -1, -1)
    return xs[0], xs[-1], ys[0], ys[-1]


# ---- geogrid: smoothing a patch without an exchange step -----------------------------------------------------------
def smoother_apron(smooth_option: int, npass: int) -> int:
    """Cells a patch must be extended by so that smoothing the extended window reproduces, on the patch itself, the
    result of smoothing the whole domain.  One 1-2-1 pass (smooth_module.F:41-55) reads one cell in every
    direction and leaves the window's outermost ring unchanged, so after p passes the values within p cells of an
    artificial window edge are off: an apron of npass cells for one_two_one (option 1), 2*npass for smth-desmth and
    smth-desmth_special (options 2, 3: each pass is a smoothing and a desmoothing sweep pair, :92-134).  The reference's
    MPI build gets the same effect with exchange_halo_r after every sweep pair."""
    return int(npass) if smooth_option == 1 else 2 * int(npass)


def apron_window(x0: int, x1: int, y0: int, y1: int, idim: int, jdim: int, apron: int) -> Tuple[int, int, int, int]:
    """The patch (1-based inclusive bounds) extended by `apron` cells, clipped to the domain: at a true domain edge the
    window's ring is the domain's ring, which the smoothers leave untouched anyway."""
    return max(1, x0 - apron), min(idim, x1 + apron), max(1, y0 - apron), min(jdim, y1 + apron)
