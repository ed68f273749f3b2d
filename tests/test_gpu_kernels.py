"""GPU numerics: every fused kernel against a plain PyTorch fp32 reference of the same op.

Multi-rank behaviour is exercised on ONE GPU through ops.local_world.LocalWorld: R virtual
ranks (own window / arena / stream each) whose kernels really handshake through flags in
device memory.  Golden model: engine.simulator.RingSimulator.
"""
import pytest
import torch

from eventgrad_b200.config import TrainConfig
from eventgrad_b200.engine.simulator import RingSimulator
from eventgrad_b200.models import build_model
from eventgrad_b200.parallel.trigger import TriggerConfig, TriggerState, trigger_step

pytestmark = pytest.mark.gpu


def _cfg(algo, **kw):
    base = dict(algo=algo, dataset="mnist", model="cnn2", lr=0.05, momentum=0.9, sync_mode="iter",
                horizon=1.0, thres_type=1, topk_percent=10.0)
    base.update(kw)
    return TrainConfig(**base).validate()


def _world(cfg, R, model="cnn2", **kw):
    from eventgrad_b200.ops.local_world import LocalWorld
    return LocalWorld(cfg, R, lambda: build_model(model), grid_cap=kw.pop("grid_cap", 6),
                      timeout_ns=5_000_000_000, **kw)


def _grads(world, n, seed, scale=0.05):
    g = torch.Generator(device="cuda").manual_seed(seed)
    out = []
    for r in range(world):
        x = torch.randn(n, generator=g, device="cuda") * scale
        out.append(x)
    return out


def _mask_pad(w, g):
    """zero the padding lanes of a flat gradient (real grads never touch them)."""
    t = w.arenas[0].table
    m = torch.zeros(t.n_padded, device="cuda")
    for o, n in zip(t.offsets, t.numels):
        m[o:o + n] = 1
    return [x * m for x in g]


@pytest.mark.parametrize("R", [1, 2, 3, 4])
@pytest.mark.parametrize("mu", [0.0, 0.9])
@pytest.mark.parametrize("dbuf", [None, False])       # None = default: double-buffered inboxes, no WAR ack
def test_decent_bitwise_vs_simulator(R, mu, dbuf):
    cfg = _cfg("decent", momentum=mu, double_buffer=dbuf)
    w = _world(cfg, R)
    assert all(be.dbuf == (dbuf is None) for be in w.backends)
    if R == 2:
        assert all(be.wire_dedup for be in w.backends)     # 2-rank ring: theta crosses the link once
    t = w.arenas[0].table
    sim = RingSimulator(R, w.arenas[0].theta.cpu(), t, "decent", lr=cfg.lr, momentum=mu, serial_skip=False)
    for s in range(7):      # odd and even inbox slots both used several times
        g = _mask_pad(w, _grads(R, t.n_padded, 100 + s))
        w.step(g)
        sim.step([x.cpu() for x in g])
    torch.cuda.synchronize()
    for be in w.backends:
        be.check_status()
    for r in range(R):
        assert torch.equal(w.arenas[r].theta.cpu(), sim.theta[r]), f"rank {r} theta differs"
        assert float(w.arenas[r].grad.abs().max()) == 0.0        # fused zero_grad
    w.close()


@pytest.mark.parametrize("R", [2, 4])
@pytest.mark.parametrize("thres", [(1, 1.0), (1, 0.9), (0, 0.0), (0, 1e9)])
def test_event_vs_simulator(R, thres):
    tt, val = thres
    cfg = _cfg("event", thres_type=tt, horizon=val, constant=val, initial_comm_passes=5)
    w = _world(cfg, R)
    t = w.arenas[0].table
    tc = TriggerConfig.from_train(cfg)
    sim = RingSimulator(R, w.arenas[0].theta.cpu(), t, "event", tc, lr=cfg.lr, momentum=cfg.momentum)
    steps = 25
    for s in range(steps):
        fires = [be.fire.clone().bool() for be in w.backends]      # decisions the kernel will act on
        g = _mask_pad(w, _grads(R, t.n_padded, 7 + s))
        w.step(g)
        sim.step([x.cpu() for x in g], fires=fires)
    torch.cuda.synchronize()
    for r, be in enumerate(w.backends):
        be.check_status()
        assert torch.equal(w.arenas[r].theta.cpu(), sim.theta[r]), f"rank {r}"
        assert be.num_events() == sim.events[r]
        assert be.bytes_sent() == sim.bytes[r]
    if tt == 0 and val == 0.0:
        assert sum(sim.events) == 2 * t.n_tensors * steps * R      # threshold 0 == dense D-PSGD
    if tt == 0 and val == 1e9:
        assert sum(sim.events) == 2 * t.n_tensors * 4 * R          # only the forced warm-up sends
    w.close()


def test_event_decisions_match_oracle():
    """Free-running kernel FSM (norm-on-write + device trigger) vs the PyTorch oracle FSM."""
    R = 2
    cfg = _cfg("event", horizon=0.95, initial_comm_passes=5)
    w = _world(cfg, R)
    t = w.arenas[0].table
    sim = RingSimulator(R, w.arenas[0].theta.cpu(), t, "event", TriggerConfig.from_train(cfg), lr=cfg.lr,
                        momentum=cfg.momentum)
    mism = tot = 0
    for s in range(30):
        fires = [be.fire.clone().bool().cpu() for be in w.backends]
        knorms = [be.cur_norm.clone() for be in w.backends]          # norm-on-write results
        for r in range(R):                                            # ... which must be accurate
            torch.testing.assert_close(knorms[r].cpu(), sim._norms(sim.theta[r]), rtol=3e-6, atol=1e-7)
        g = _mask_pad(w, _grads(R, t.n_padded, 500 + s))
        w.step(g)
        sim.step([x.cpu() for x in g], norms=knorms)
        for r in range(R):
            mism += int((fires[r] != sim.fire_history[-1][r]).sum())
            tot += t.n_tensors
        if mism:
            break      # trajectories diverge after the first differing decision
    torch.cuda.synchronize()
    assert mism == 0, f"{mism}/{tot} trigger decisions differ from the oracle"
    for r in range(R):
        torch.testing.assert_close(w.arenas[r].theta.cpu(), sim.theta[r], rtol=0, atol=0)
    w.close()


def test_fsm_kernel_vs_oracle_random_norms():
    from eventgrad_b200.ops.local_world import LocalWorld
    cfg = _cfg("event", horizon=0.9, initial_comm_passes=3)
    w = _world(cfg, 1)
    be = w.backends[0]
    sz = w.arenas[0].table.n_tensors
    tc = TriggerConfig.from_train(cfg)
    st = TriggerState(sz, cfg.sent_history)
    # reset device FSM to a clean state
    for x in (be.thres, be.last_norm, be.last_iter, be.slopes):
        x.zero_()
    be.d_pass.zero_()
    g = torch.Generator().manual_seed(0)
    norms = torch.rand(sz, generator=g) + 1.0
    for k in range(1, 40):
        norms = norms + (torch.rand(sz, generator=g) - 0.45) * 0.01 * (1 + (k % 7 == 0) * 5)
        dn = norms.float().cuda()
        be.C.fsm_decide(be.gp, dn.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        fire = trigger_step(st, norms.float(), k, tc)
        assert torch.equal(be.fire.cpu().bool(), fire), f"step {k}"
        assert torch.equal(be.thres.cpu(), st.thres)
        assert torch.equal(be.last_norm.cpu(), st.last_norm)
        assert torch.equal(be.slopes.cpu().view(sz, -1), st.slopes)
    w.close()


@pytest.mark.parametrize("R", [2, 3])
@pytest.mark.parametrize("pct", [1.0, 10.0, 100.0])
def test_spevent_vs_simulator(R, pct):
    cfg = _cfg("spevent", topk_percent=pct, initial_comm_passes=4, horizon=1.0)
    w = _world(cfg, R)
    t = w.arenas[0].table
    sim = RingSimulator(R, w.arenas[0].theta.cpu(), t, "spevent", TriggerConfig.from_train(cfg), lr=cfg.lr,
                        momentum=cfg.momentum, topk_percent=pct)
    for s in range(12):
        fires = [be.fire.clone().bool() for be in w.backends]
        g = _mask_pad(w, _grads(R, t.n_padded, 900 + s))
        w.step(g)
        sim.step([x.cpu() for x in g], fires=fires)
    torch.cuda.synchronize()
    for r, be in enumerate(w.backends):
        be.check_status()
        assert torch.equal(be.prev.cpu(), sim.prev[r]), f"prev rank {r}"
        assert torch.equal(be.rep_l.cpu(), sim.rep_l[r]), f"rep_l rank {r}"
        assert torch.equal(be.rep_r.cpu(), sim.rep_r[r]), f"rep_r rank {r}"
        assert torch.equal(w.arenas[r].theta.cpu(), sim.theta[r]), f"theta rank {r}"
        assert be.num_events() == sim.events[r]
        assert be.bytes_sent() == sim.bytes[r]
    w.close()


@pytest.mark.parametrize("R", [2, 3])
@pytest.mark.parametrize("pct", [1.0, 10.0, 37.0])
def test_spevent_heavy_ties_resolved_towards_lowest_index(R, pct):
    """Three-valued gradients => thousands of EQUAL |theta - prev| keys per tensor, so the k-th largest key sits inside
    a big tie group spanning many tiles: exercises the exact-threshold select (candidate bucket = almost everything),
    the tie look-back across tiles and the 'lowest index wins' rule -- bitwise vs the oracle (stable sort)."""
    cfg = _cfg("spevent", topk_percent=pct, initial_comm_passes=3, horizon=1.0, momentum=0.0)
    w = _world(cfg, R)
    t = w.arenas[0].table
    sim = RingSimulator(R, w.arenas[0].theta.cpu(), t, "spevent", TriggerConfig.from_train(cfg), lr=cfg.lr,
                        momentum=0.0, topk_percent=pct)
    gen = torch.Generator(device="cuda").manual_seed(5)
    for s in range(8):
        fires = [be.fire.clone().bool() for be in w.backends]
        g = []
        for r in range(R):
            u = torch.rand(t.n_padded, generator=gen, device="cuda")
            g.append(torch.where(u < 0.3, -0.0625, torch.where(u < 0.6, 0.0, 0.0625)).to(torch.float32))
        g = _mask_pad(w, g)
        w.step(g)
        sim.step([x.cpu() for x in g], fires=fires)
    torch.cuda.synchronize()
    for r, be in enumerate(w.backends):
        be.check_status()
        assert torch.equal(be.prev.cpu(), sim.prev[r]), f"prev rank {r}"
        assert torch.equal(be.rep_l.cpu(), sim.rep_l[r]), f"rep_l rank {r}"
        assert torch.equal(be.rep_r.cpu(), sim.rep_r[r]), f"rep_r rank {r}"
        assert torch.equal(w.arenas[r].theta.cpu(), sim.theta[r]), f"theta rank {r}"
    w.close()


@pytest.mark.parametrize("R", [2, 4])
@pytest.mark.parametrize("model", ["mlp", "resnet18"])
def test_cent_allreduce_vs_reference(R, model):
    cfg = _cfg("cent", model=model, momentum=0.0 if model == "mlp" else 0.9, lr=1e-2)
    w = _world(cfg, R, model=model, grid_cap=16)
    t = w.arenas[0].table
    theta0 = w.arenas[0].theta.clone()
    mom = torch.zeros_like(theta0)
    ref = theta0.clone()
    for s in range(3):
        g = _mask_pad(w, _grads(R, t.n_padded, 40 + s))
        w.step(g)
        acc = g[0].clone()
        for r in range(1, R):
            acc = acc + g[r]                       # rank order, fp32
        gbar = acc / float(R)
        if cfg.momentum:
            mom = mom * cfg.momentum + gbar
            ref = torch.addcmul(ref, mom, torch.tensor(-cfg.lr, device="cuda"))
        else:
            ref = torch.addcmul(ref, gbar, torch.tensor(-cfg.lr, device="cuda"))
    torch.cuda.synchronize()
    for r, be in enumerate(w.backends):
        be.check_status()
        torch.testing.assert_close(w.arenas[r].theta, ref, rtol=1e-6, atol=1e-7)
        assert torch.equal(w.arenas[r].theta, w.arenas[0].theta)       # replicas stay bit-identical
        assert float(w.arenas[r].grad.abs().max()) == 0.0
    w.close()


def test_final_average():
    R = 3
    cfg = _cfg("event")
    w = _world(cfg, R)
    n = w.arenas[0].table.n_padded
    vals = [torch.randn(n, device="cuda") for _ in range(R)]
    for r in range(R):
        w.arenas[r].theta.copy_(vals[r])
    torch.cuda.synchronize()
    w.final_average()
    ref = ((vals[0] + vals[1]) + vals[2]) / 3.0
    for r in range(R):
        torch.testing.assert_close(w.arenas[r].theta, ref, rtol=1e-6, atol=1e-7)
        assert torch.equal(w.arenas[r].theta, w.arenas[0].theta)
    w.close()


def test_async_mode_runs_and_counts():
    R = 2
    cfg = _cfg("event", sync_mode="async", thres_type=0, constant=0.0)
    w = _world(cfg, R)
    t = w.arenas[0].table
    for s in range(6):
        w.step(_mask_pad(w, _grads(R, t.n_padded, s)))
    torch.cuda.synchronize()
    for be in w.backends:
        be.check_status()
        assert be.num_events() == 2 * t.n_tensors * 6
    assert all(torch.isfinite(a.theta).all() for a in w.arenas)
    w.close()


def test_norm_on_write_matches_torch_norm():
    cfg = _cfg("event", model="resnet18", dataset="cifar10")
    w = _world(cfg, 1, model="resnet18", grid_cap=64)
    a, be = w.arenas[0], w.backends[0]
    ref = a.tensor_norms()
    torch.testing.assert_close(be.cur_norm, ref, rtol=2e-6, atol=1e-7)
    w.close()


@pytest.mark.parametrize("nhwc", [False, True])
@pytest.mark.parametrize("bf16", [False, True])
def test_decode_augment_kernel(nhwc, bf16):
    from eventgrad_b200.data.augment import decode_augment_torch, draw_augment_params
    from eventgrad_b200.ops.augment import decode_augment
    x = torch.randint(0, 256, (37, 3, 32, 32), dtype=torch.uint8, device="cuda")
    p = draw_augment_params(37, 4, "cuda")
    dt = torch.bfloat16 if bf16 else torch.float32
    out = decode_augment(x, 1 / 255.0, 0.45, 0.25, p, out_dtype=dt, channels_last=nhwc)
    ref = decode_augment_torch(x, 1 / 255.0, 0.45, 0.25, p, out_dtype=dt)
    torch.testing.assert_close(out.float(), ref.float(), rtol=1e-2 if bf16 else 1e-6, atol=1e-2 if bf16 else 1e-6)
    if nhwc:
        assert out.is_contiguous(memory_format=torch.channels_last)
    out2 = decode_augment(x, 1.0, 0.0, 1.0, None)
    assert torch.equal(out2, x.float())


@pytest.mark.parametrize("algo", ["decent", "event", "spevent"])
@pytest.mark.parametrize("R", [2, 3])
def test_split_step_overlap_mode_matches_simulator(algo, R):
    """overlap_push: push half (side stream in the trainer) + wait/mix half == the fused kernel."""
    cfg = _cfg(algo, overlap_push=True, initial_comm_passes=4, horizon=1.0)
    w = _world(cfg, R)
    assert all(be.overlap for be in w.backends)
    t = w.arenas[0].table
    sim = RingSimulator(R, w.arenas[0].theta.cpu(), t, algo, TriggerConfig.from_train(cfg), lr=cfg.lr,
                        momentum=cfg.momentum, topk_percent=cfg.topk_percent, serial_skip=False)
    for s in range(10):
        fires = [be.fire.clone().bool() for be in w.backends] if algo != "decent" else None
        g = _mask_pad(w, _grads(R, t.n_padded, 321 + s))
        w.step(g)
        sim.step([x.cpu() for x in g], fires=fires)
    torch.cuda.synchronize()
    for r, be in enumerate(w.backends):
        be.check_status()
        assert torch.equal(w.arenas[r].theta.cpu(), sim.theta[r]), f"rank {r}"
        if algo != "decent":
            assert be.num_events() == sim.events[r]
    w.close()


@pytest.mark.parametrize("R", [1, 2, 3])
def test_ce_push_split_step_matches_simulator(R):
    """csrc/ce_push.cu: ack wait kernel -> cudaMemcpyAsync (x1 on a 2-rank ring, else x2) -> pushed-flag kernel as the
    push half of the split step (decent), R virtual ranks on one GPU, bit-exact vs the oracle."""
    cfg = _cfg("decent", overlap_push=True, ce_push=True)
    w = _world(cfg, R)
    assert all(be.ce_push for be in w.backends)
    t = w.arenas[0].table
    sim = RingSimulator(R, w.arenas[0].theta.cpu(), t, "decent", lr=cfg.lr, momentum=cfg.momentum, serial_skip=False)
    for s in range(8):
        g = _mask_pad(w, _grads(R, t.n_padded, 700 + s))
        w.step(g)
        sim.step([x.cpu() for x in g])
    torch.cuda.synchronize()
    for r, be in enumerate(w.backends):
        be.check_status()
        assert torch.equal(w.arenas[r].theta.cpu(), sim.theta[r]), f"rank {r}"
    w.close()


@pytest.mark.parametrize("mode", ["fused_ack", "fused_dbuf", "split"])
def test_dead_peer_trips_timeout_and_stops_pushing(mode):
    """Failure detection on the PRODUCT backend: virtual rank 1 stops stepping; rank 0's next step must come back
    within the device timeout with the sticky status EG_ERR_TIMEOUT (no hang), and from then on rank 0 must not
    store into the dead peer's inbox any more (honour s_ok)."""
    import time
    from eventgrad_b200.ops.local_world import LocalWorld
    kw = {"fused_ack": dict(double_buffer=False), "fused_dbuf": dict(), "split": dict(overlap_push=True)}[mode]
    cfg = _cfg("decent", **kw)
    w = LocalWorld(cfg, 2, lambda: build_model("cnn2"), grid_cap=6, timeout_ns=300_000_000)     # 0.3 s
    t = w.arenas[0].table
    for s in range(3):
        w.step(_mask_pad(w, _grads(2, t.n_padded, 900 + s)))
    torch.cuda.synchronize()
    for be in w.backends:
        be.check_status()
    peer_inbox = w.backends[1].win.view("inbox_r", torch.float32)
    t0 = time.perf_counter()
    with torch.cuda.stream(w.streams[0]):          # only rank 0 steps: its neighbour is "dead"
        w.backends[0].step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert dt < 5.0, f"step against a dead peer took {dt:.1f} s (timeout is 0.3 s per wait)"
    with pytest.raises(RuntimeError, match="status 1"):
        w.backends[0].check_status()
    snap = peer_inbox.clone()
    t0 = time.perf_counter()
    with torch.cuda.stream(w.streams[0]):          # status is sticky: waits return at once, nothing is stored
        w.backends[0].step()
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 2.0
    assert torch.equal(peer_inbox, snap), "rank 0 kept writing into the dead peer's inbox after the timeout"
    w.close()
