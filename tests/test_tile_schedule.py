"""Host logic of the tiled kernel's work order (sn_tile_schedule, no GPU): the order replaces the barrier between colour
passes of the reference-free design, so its invariants are what makes the chain a valid sequential sweep:
every tile once per sweep, tiles that are active together never adjacent (also across the periodic wrap, with partial
last tiles and odd tile counts), phases in order."""
import itertools

import numpy as np
import pytest


@pytest.fixture(scope="module")
def sn(built):
    import starrynight_b200
    return starrynight_b200


SHAPES = [(32, 32, 32), (64, 32, 48), (48, 48, 48), (100, 44, 36), (20, 27, 24), (33, 50, 40), (112, 32, 144), (512, 512, 64), (250, 250, 252)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("reps", [1, 3])
def test_every_tile_once_and_neighbours_never_together(sn, shape, reps):
    X, Y, Z = shape
    tn = [(n + 15) // 16 for n in shape]
    items = sn.tile_schedule(X, Y, Z, reps, sweep=5)
    assert len(items) == reps * tn[0] * tn[1] * tn[2]
    seen = set(map(tuple, items[:, :4].tolist()))
    assert len(seen) == len(items), "a tile is visited twice in one sweep"
    assert seen == set(itertools.product(range(reps), range(tn[0]), range(tn[1]), range(tn[2])))
    phase = {tuple(r[:4]): int(r[4]) for r in items.tolist()}
    ncol = [3 if t % 2 else 2 for t in tn]
    assert items[:, 4].max() + 1 == ncol[0] * ncol[1] * ncol[2]
    assert np.all(np.diff(items[:, 4]) >= 0), "phases are not visited in order"
    # tiles of one phase are mutually non-adjacent (26-neighbourhood, periodic): their 22^3 read sets miss each other's 16^3 write sets
    for (r, tx, ty, tz), p in phase.items():
        if r:
            continue
        for dx, dy, dz in itertools.product((-1, 0, 1), repeat=3):
            n = ((tx + dx) % tn[0], (ty + dy) % tn[1], (tz + dz) % tn[2])
            if n != (tx, ty, tz):
                assert phase[(0,) + n] != p, f"tiles {(tx, ty, tz)} and {n} are neighbours and share phase {p}"


def test_the_phase_of_a_tile_does_not_depend_on_the_sweep_but_the_order_inside_a_phase_rotates(sn):
    a, b = sn.tile_schedule(128, 64, 64, 1, sweep=0), sn.tile_schedule(128, 64, 64, 1, sweep=1)
    pa = {tuple(r[:4]): int(r[4]) for r in a.tolist()}
    pb = {tuple(r[:4]): int(r[4]) for r in b.tolist()}
    assert pa == pb
    assert not np.array_equal(a, b), "the x order is rotated by one tile plane per sweep"
    # replicas of one phase follow each other: a phase is (replica, tile) in that order
    c = sn.tile_schedule(64, 64, 64, 2, sweep=0)
    first = c[c[:, 4] == 0]
    assert list(first[:, 0]) == sorted(first[:, 0])


def test_rejects_what_the_tiled_kernel_cannot_take(sn):
    with pytest.raises(sn.SnError):
        sn.tile_schedule(16, 64, 64)
    with pytest.raises(sn.SnError):
        sn.tile_schedule(64, 64, 30)


@pytest.mark.parametrize("shape", [(64, 32, 48), (48, 48, 48), (100, 44, 36), (20, 27, 24)])
def test_every_dependency_precedes_its_item_in_the_global_order(sn, shape):
    """The dataflow rule (sn_tile_deps_ready): an item (sweep s, phase p) waits until each neighbouring tile has completed
    s + 1 sweeps if its phase q < p, else s sweeps.  Every such event must lie earlier in the global item order, or persistent
    CTAs that take items in order could wait for work nobody has started (deadlock)."""
    X, Y, Z = shape
    tn = [(n + 15) // 16 for n in shape]
    per_sweep = []
    for s in range(3):
        items = sn.tile_schedule(X, Y, Z, 1, sweep=s)
        per_sweep.append({tuple(r[1:4]): (i, int(r[4])) for i, r in enumerate(items.tolist())})     # tile -> (position, phase)
    S = len(per_sweep[0])
    for s in (1, 2):
        for tile, (pos, p) in per_sweep[s].items():
            n_item = s * S + pos
            for d in itertools.product((-1, 0, 1), repeat=3):
                nb = tuple((tile[a] + d[a]) % tn[a] for a in range(3))
                if nb == tile:
                    continue
                q = per_sweep[s][nb][1]
                need_sweep = s if q < p else s - 1                   # the neighbour's item whose completion is awaited
                n_dep = need_sweep * S + per_sweep[need_sweep][nb][0]
                assert n_dep < n_item, f"tile {tile} (sweep {s}, phase {p}) waits for {nb} (phase {q}) which comes later in the order"
