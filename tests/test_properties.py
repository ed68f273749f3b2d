"""Algorithm-level invariants of the golden model (which the kernels reproduce bit-for-bit)."""
import torch
from hypothesis import given, settings, strategies as st

from eventgrad_b200.engine.simulator import RingSimulator
from eventgrad_b200.models import build_model
from eventgrad_b200.parallel.arena import ParamArena
from eventgrad_b200.parallel.trigger import TriggerConfig

torch.manual_seed(0)
_ARENA = ParamArena(build_model("cnn2"))
_T = _ARENA.table
_MASK = torch.zeros(_T.n_padded)
for _o, _n in zip(_T.offsets, _T.numels):
    _MASK[_o:_o + _n] = 1


def _rand_thetas(R, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(_T.n_padded, generator=g) * _MASK for _ in range(R)]


@settings(max_examples=12, deadline=None)
@given(st.integers(2, 7), st.integers(0, 10_000))
def test_dense_gossip_preserves_the_network_average_and_contracts(R, seed):
    """W = (I + P + P^T)/3 is doubly stochastic (also for R = 2, where the peer is counted twice): with
    lr = 0 the average over ranks is invariant and the disagreement shrinks every round."""
    sim = RingSimulator(R, _ARENA.theta.clone(), _T, "decent", lr=0.0, serial_skip=False)
    sim.theta = _rand_thetas(R, seed)
    mean0 = torch.stack(sim.theta).mean(0)
    dis_prev = float(torch.stack(sim.theta).var(0, unbiased=False).sum())
    zeros = [torch.zeros(_T.n_padded)] * R
    for _ in range(6):
        sim.step(zeros)
        th = torch.stack(sim.theta)
        assert torch.allclose(th.mean(0), mean0, atol=1e-5)
        dis = float(th.var(0, unbiased=False).sum())
        assert dis <= dis_prev * (1 + 1e-6)
        dis_prev = dis
    if R in (2, 3):          # spectral gap 1/3 (R=2) and complete averaging (R=3) -> essentially consensus
        assert dis_prev < 1e-3 * max(1.0, float(mean0.abs().sum()))


@settings(max_examples=10, deadline=None)
@given(st.integers(2, 5), st.sampled_from([1.0, 10.0, 100.0]), st.integers(0, 1000))
def test_sparse_bookkeeping_invariants(R, pct, seed):
    """After a send, prev equals theta exactly at the sent indices, every neighbour's replica of me equals
    my prev, and bytes = sum over fired tensors of 2 neighbours x 2 k_i words."""
    tc = TriggerConfig(1, 1.0, 0.0, 2, 3)
    sim = RingSimulator(R, _ARENA.theta.clone(), _T, "spevent", tc, lr=0.05, topk_percent=pct, serial_skip=False)
    g = torch.Generator().manual_seed(seed)
    for s in range(5):
        sim.step([torch.randn(_T.n_padded, generator=g) * 0.05 * _MASK for _ in range(R)])
    for r in range(R):
        L, Rn = (r - 1) % R, (r + 1) % R
        assert torch.equal(sim.rep_r[L], sim.prev[r])          # left neighbour's view of me
        assert torch.equal(sim.rep_l[Rn], sim.prev[r])         # right neighbour's view of me
    k = _T.topk_counts(pct)
    want = [sum(2 * 2 * k[i] * 4 for step in sim.fire_history for i in range(_T.n_tensors) if step[r][i])
            for r in range(R)]
    assert sim.bytes == want
    assert all(e % 2 == 0 for e in sim.events)


def test_event_with_huge_threshold_after_warmup_freezes_inboxes():
    """constant = +inf: after the forced warm-up nothing is ever sent again, so every rank keeps mixing
    with the frozen copies it received last (the reference's behaviour when a neighbour goes silent)."""
    sim = RingSimulator(3, _ARENA.theta.clone(), _T, "event", TriggerConfig(0, 1.0, float("inf"), 2, 4), lr=0.05)
    g = torch.Generator().manual_seed(0)
    for s in range(3):
        sim.step([torch.randn(_T.n_padded, generator=g) * 0.05 * _MASK for _ in range(3)])
    frozen = [b.clone() for b in sim.inbox_l]
    ev = sim.total_events()
    for s in range(4):
        sim.step([torch.randn(_T.n_padded, generator=g) * 0.05 * _MASK for _ in range(3)])
    assert sim.total_events() == ev
    assert all(torch.equal(a, b) for a, b in zip(frozen, sim.inbox_l))
