"""Property tests of the trigger FSM oracle."""
import math

import torch
from hypothesis import given, settings, strategies as st

from eventgrad_b200.parallel.trigger import TriggerConfig, TriggerState, trigger_step, mix3_, sgd_, topk_select


def _scalar_reference(norms_seq, cfg):
    """Straight transcription of the per-tensor scalar logic (event.cpp:300-355) for ONE tensor."""
    thres = last_norm = last_iter = 0.0
    slopes = [0.0] * cfg.sent_history
    fired = []
    f32 = lambda x: float(torch.tensor(x, dtype=torch.float32))
    for k, cur in enumerate(norms_seq, start=1):
        cur = f32(cur)
        vd = f32(abs(f32(cur - last_norm)))
        it = f32(k - last_iter)
        thres = f32(thres * f32(cfg.horizon)) if cfg.thres_type == 1 else f32(cfg.constant)
        fire = vd >= thres or k < cfg.initial_comm_passes
        if fire:
            slopes = slopes[1:] + [f32(vd / it)]
            avg = sum(slopes) / cfg.sent_history
            if cfg.thres_type == 1:
                thres = f32(avg)
            last_norm, last_iter = cur, float(k)
        fired.append(fire)
    return fired


@settings(max_examples=60, deadline=None)
@given(st.lists(st.floats(0.1, 10.0), min_size=5, max_size=60),
       st.sampled_from([0.0, 0.5, 0.9, 1.0]), st.integers(0, 10))
def test_vectorised_oracle_matches_scalar_transcription(norms, horizon, warm):
    cfg = TriggerConfig(1, horizon, 0.0, 2, warm)
    stt = TriggerState(1, 2)
    got = [bool(trigger_step(stt, torch.tensor([n], dtype=torch.float32), k, cfg)[0])
           for k, n in enumerate(norms, start=1)]
    assert got == _scalar_reference(norms, cfg)


@settings(max_examples=30, deadline=None)
@given(st.lists(st.floats(0.1, 10.0), min_size=35, max_size=50))
def test_forced_sends_during_warmup_and_zero_threshold(norms):
    cfg = TriggerConfig(1, 1.0, 0.0, 2, 30)
    stt = TriggerState(3, 2)
    for k, n in enumerate(norms, start=1):
        f = trigger_step(stt, torch.full((3,), n, dtype=torch.float32), k, cfg)
        if k < 30:
            assert f.all()                               # pass_num < initial_comm_passes
    for tt, hz, c in ((1, 0.0, 0.0), (0, 1.0, 0.0)):     # horizon 0 / constant 0 => always fire
        cfg0 = TriggerConfig(tt, hz, c, 2, 0)
        s0 = TriggerState(2, 2)
        for k, n in enumerate(norms, start=1):
            assert trigger_step(s0, torch.full((2,), n, dtype=torch.float32), k, cfg0).all()


def test_infinite_constant_only_warmup():
    cfg = TriggerConfig(0, 1.0, float("inf"), 2, 5)
    stt = TriggerState(4, 2)
    n = 0
    for k in range(1, 40):
        n += int(trigger_step(stt, torch.rand(4) + k, k, cfg).sum())
    assert n == 4 * 4


def test_mix_and_sgd_semantics():
    t, l, r = torch.tensor([3.0]), torch.tensor([6.0]), torch.tensor([0.0])
    assert mix3_(t, l, r).item() == 3.0
    th, g, m = torch.tensor([1.0]), torch.tensor([2.0]), torch.tensor([0.0])
    sgd_(th, g, m, 0.1, 0.9)
    assert math.isclose(th.item(), 0.8, rel_tol=1e-6) and m.item() == 2.0     # first step: buf = g
    sgd_(th, g, m, 0.1, 0.9)
    assert math.isclose(m.item(), 3.8, rel_tol=1e-6)


def test_topk_tie_break_lowest_index():
    th, pv = torch.tensor([1.0, 1.0, 1.0, 5.0, 1.0]), torch.zeros(5)
    v, ix = topk_select(th, pv, 3)
    assert ix.tolist() == [3, 0, 1] and v.tolist() == [5.0, 1.0, 1.0]
