"""Randomised-interleaving model check of the iter-sync exchange protocol (PROTOCOL.md section 1).

Every warp of every rank is a little state machine that performs, in program order, exactly the
synchronisation actions of `gossip_step_kernel` (ack wait, push, flag publish, flag wait, mix, ticket,
ack).  A random scheduler interleaves all warps of all ranks; blocking conditions are the `wait_ge`
spins.  Checked: (1) no deadlock, (2) every mix of step k reads, for a slice the neighbour fired at step
k, exactly version k, and otherwise the neighbour's LAST fired version <= k-1 -- never a value from the
future (the WAR hazard the ack guards against).  A mutant without the ack wait must be caught.
"""
import random

import pytest


class Rank:
    def __init__(self, n_tiles, n_warps):
        self.inbox = {"l": [[0] * n_warps for _ in range(n_tiles)], "r": [[0] * n_warps for _ in range(n_tiles)]}
        self.flag = {"l": [[0] * n_warps for _ in range(n_tiles)], "r": [[0] * n_warps for _ in range(n_tiles)]}
        self.ack = {"l": 0, "r": 0}
        self.ticket = 0
        self.pass_num = 0


def run(R, G, n_tiles, n_warps, steps, depth, seed, use_ack=True, max_ticks=400000):
    rng = random.Random(seed)
    ranks = [Rank(n_tiles, n_warps) for _ in range(R)]
    fire = {(r, t, k): rng.random() < 0.6 for r in range(R) for t in range(n_tiles) for k in range(1, steps + 1)}
    last_fired = {}                                  # (rank, tile) -> list of steps fired so far (by program order)
    violations = []
    iters = -(-n_tiles // G)

    def warp(r, b, w):
        L, Rn = (r - 1) % R, (r + 1) % R
        me = ranks[r]
        for k in range(1, steps + 1):
            if use_ack:
                yield lambda: me.ack["l"] >= k - 1 and me.ack["r"] >= k - 1       # WAR guard
            for j in range(iters + depth):
                t = b + j * G
                if j < iters and t < n_tiles:
                    if fire[(r, t, k)]:
                        ranks[L].inbox["r"][t][w] = k                              # push my slice to both neighbours
                        ranks[Rn].inbox["l"][t][w] = k
                        yield None
                    ranks[L].flag["r"][t][w] = k                                    # publish (release) flags
                    ranks[Rn].flag["l"][t][w] = k
                    yield None
                t2 = b + (j - depth) * G
                if j >= depth and t2 < n_tiles:
                    yield lambda t2=t2: me.flag["l"][t2][w] >= k and me.flag["r"][t2][w] >= k
                    for side, nb in (("l", L), ("r", Rn)):                          # mix: read the inbox slice
                        got = me.inbox[side][t2][w]
                        want = max([s for s in range(1, k + 1) if fire[(nb, t2, s)]] or [0])
                        if got != want:
                            violations.append((r, side, t2, w, k, got, want))
                    yield None
            # CTA/grid tail: modelled per warp -- the last warp of the rank to finish publishes the acks
            me.ticket += 1
            if me.ticket == G * n_warps * k:
                me.pass_num = k
                ranks[L].ack["r"] = k
                ranks[Rn].ack["l"] = k
            yield None

    actors = [warp(r, b, w) for r in range(R) for b in range(G) for w in range(n_warps)]
    pending = [None] * len(actors)                   # blocking condition each actor is currently waiting on
    alive = set(range(len(actors)))
    for i in list(alive):
        pending[i] = next(actors[i])
    ticks = 0
    while alive and ticks < max_ticks:
        ready = [i for i in alive if pending[i] is None or pending[i]()]
        if not ready:
            return "deadlock", violations
        i = rng.choice(ready)
        try:
            pending[i] = next(actors[i])
        except StopIteration:
            alive.discard(i)
        ticks += 1
    return ("ok" if not alive else "timeout"), violations


@pytest.mark.parametrize("R,G,n_tiles,n_warps,depth", [(2, 2, 5, 2, 2), (3, 2, 5, 2, 2), (4, 3, 7, 2, 1), (3, 1, 4, 3, 4),
                                                      (1, 2, 3, 2, 2)])
def test_protocol_is_deadlock_free_and_never_reads_the_future(R, G, n_tiles, n_warps, depth):
    for seed in range(40):
        status, viol = run(R, G, n_tiles, n_warps, steps=4, depth=depth, seed=seed)
        assert status == "ok", (status, seed)
        assert not viol, viol[:3]


def test_mutant_without_ack_is_caught():
    """Remove the WAR guard: some schedule lets a fast neighbour overwrite an inbox slice with step k+1
    before it was mixed at step k -- the checker must see it."""
    caught = 0
    for seed in range(60):
        status, viol = run(3, 2, 5, 2, steps=4, depth=2, seed=seed, use_ack=False)
        assert status == "ok"
        caught += bool(viol)
    assert caught > 0


def run_double_buffered(R, G, n_tiles, n_warps, steps, depth, seed, max_ticks=400000):
    """NEXT_STEPS.md item 2: dense gossip (every slice fires every step) with TWO inbox slots
    (slot = step & 1) and NO ack at all."""
    rng = random.Random(seed)
    inbox = [{side: [[[0, 0] for _ in range(n_warps)] for _ in range(n_tiles)] for side in "lr"} for _ in range(R)]
    flag = [{side: [[0] * n_warps for _ in range(n_tiles)] for side in "lr"} for _ in range(R)]
    violations = []
    iters = -(-n_tiles // G)

    def warp(r, b, w):
        L, Rn = (r - 1) % R, (r + 1) % R
        for k in range(1, steps + 1):
            for j in range(iters + depth):
                t = b + j * G
                if j < iters and t < n_tiles:
                    inbox[L]["r"][t][w][k & 1] = k
                    inbox[Rn]["l"][t][w][k & 1] = k
                    yield None
                    flag[L]["r"][t][w] = k
                    flag[Rn]["l"][t][w] = k
                    yield None
                t2 = b + (j - depth) * G
                if j >= depth and t2 < n_tiles:
                    yield lambda t2=t2: flag[r]["l"][t2][w] >= k and flag[r]["r"][t2][w] >= k
                    for side in "lr":
                        got = inbox[r][side][t2][w][k & 1]
                        if got != k:
                            violations.append((r, side, t2, w, k, got))
                    yield None

    actors = [warp(r, b, w) for r in range(R) for b in range(G) for w in range(n_warps)]
    pending = [next(a) for a in actors]
    alive = set(range(len(actors)))
    ticks = 0
    while alive and ticks < max_ticks:
        ready = [i for i in alive if pending[i] is None or pending[i]()]
        if not ready:
            return "deadlock", violations
        i = rng.choice(ready)
        try:
            pending[i] = next(actors[i])
        except StopIteration:
            alive.discard(i)
        ticks += 1
    return ("ok" if not alive else "timeout"), violations


@pytest.mark.parametrize("R,G,n_tiles,n_warps,depth", [(2, 2, 5, 2, 2), (3, 2, 5, 2, 2), (4, 2, 6, 2, 1), (5, 1, 4, 2, 3)])
def test_ack_free_double_buffered_dense_gossip_design(R, G, n_tiles, n_warps, depth):
    """Design check for the next round: with slot = step & 1 a sender can be at most one step ahead of a
    receiver's reads (it needs the receiver's flags of step k+1 to finish k+1), so no ack is required."""
    for seed in range(40):
        status, viol = run_double_buffered(R, G, n_tiles, n_warps, steps=5, depth=depth, seed=seed)
        assert status == "ok", (status, seed)
        assert not viol, viol[:3]


def run_split_step(R, n_ctas, n_tiles, steps, seed, use_ack=True, fire_p=0.6, max_ticks=400000):
    """PROTOCOL.md section 2 (also the copy-engine variant csrc/ce_push.cu, which keeps the same flags):
    phase 1 = [ack wait] -> push every fired tile -> publish pushed=k; phase 2 (after phase 1 of the same step in
    stream order) = wait pushed_from_{l,r} >= k -> every CTA mixes its tiles -> last CTA acks k.  Phase 1 of
    step k+1 starts after phase 2 of step k (side.wait_stream(cur))."""
    rng = random.Random(seed)
    inbox = [{s: [0] * n_tiles for s in "lr"} for _ in range(R)]
    pushed = [{s: 0 for s in "lr"} for _ in range(R)]
    ack = [{s: 0 for s in "lr"} for _ in range(R)]
    done2 = [0] * R                                   # CTAs of rank r that finished phase 2 (cumulative)
    p1_done = [0] * R                                 # last step whose phase 1 completed on rank r
    fire = {(r, t, k): rng.random() < fire_p for r in range(R) for t in range(n_tiles) for k in range(1, steps + 1)}
    violations = []

    def phase1(r):
        L, Rn = (r - 1) % R, (r + 1) % R
        for k in range(1, steps + 1):
            yield lambda k=k: done2[r] >= n_ctas * (k - 1)                     # stream order: after phase 2 of k-1
            if use_ack:
                yield lambda k=k: ack[r]["l"] >= k - 1 and ack[r]["r"] >= k - 1
            for t in range(n_tiles):
                if fire[(r, t, k)]:
                    inbox[L]["r"][t] = k
                    inbox[Rn]["l"][t] = k
                    yield None
            pushed[L]["r"] = k
            pushed[Rn]["l"] = k
            p1_done[r] = k
            yield None

    def phase2(r, b):
        L, Rn = (r - 1) % R, (r + 1) % R
        for k in range(1, steps + 1):
            yield lambda k=k: p1_done[r] >= k and done2[r] >= n_ctas * (k - 1)    # stream order
            yield lambda k=k: pushed[r]["l"] >= k and pushed[r]["r"] >= k
            for t in range(b, n_tiles, n_ctas):
                for side, nb in (("l", L), ("r", Rn)):
                    got = inbox[r][side][t]
                    want = max([s for s in range(1, k + 1) if fire[(nb, t, s)]] or [0])
                    if got != want:
                        violations.append((r, side, t, k, got, want))
                yield None
            done2[r] += 1
            if done2[r] == n_ctas * k:                                          # last CTA of the grid acks
                ack[L]["r"] = k
                ack[Rn]["l"] = k
            yield None

    actors = [phase1(r) for r in range(R)] + [phase2(r, b) for r in range(R) for b in range(n_ctas)]
    pending = [next(a) for a in actors]
    alive = set(range(len(actors)))
    ticks = 0
    while alive and ticks < max_ticks:
        ready = [i for i in alive if pending[i] is None or pending[i]()]
        if not ready:
            return "deadlock", violations
        i = rng.choice(ready)
        try:
            pending[i] = next(actors[i])
        except StopIteration:
            alive.discard(i)
        ticks += 1
    return ("ok" if not alive else "timeout"), violations


@pytest.mark.parametrize("R,n_ctas,n_tiles", [(1, 2, 4), (2, 2, 5), (3, 3, 7), (5, 2, 4)])
def test_split_step_protocol(R, n_ctas, n_tiles):
    for seed in range(40):
        status, viol = run_split_step(R, n_ctas, n_tiles, steps=5, seed=seed)
        assert status == "ok", (status, seed)
        assert not viol, viol[:3]


def test_split_step_mutant_without_ack_is_caught():
    caught = 0
    for seed in range(80):
        status, viol = run_split_step(3, 2, 5, steps=5, seed=seed, use_ack=False)
        assert status == "ok"
        caught += bool(viol)
    assert caught > 0
