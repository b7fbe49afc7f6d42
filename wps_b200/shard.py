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
