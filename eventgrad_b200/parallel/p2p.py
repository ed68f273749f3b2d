"""Fused peer-to-peer backend -- the product path.

Per training step, after backward():
  cent     ONE launch: one-/two-shot all-reduce over peer-mapped gradients fused with 1/R and SGD
  decent   ONE launch: push theta to both ring neighbours -> flag handshake -> (t+L+R)/3 -> SGD
  event    ONE launch: as decent, but a tensor is pushed only if its device-resident trigger fired;
           the same kernel accumulates the norms and runs the trigger FSM for the next step
  spevent  batched segmented radix-select top-k + compaction straight into the neighbours'
           inboxes, scatter of arrived records into replicas, then the dense mix+SGD kernel
No NCCL/MPI call, no separate elementwise kernel and no host synchronisation on that path
(BASELINE.json north star).  Kernels: csrc/gossip.cu, csrc/allreduce.cu, csrc/sparse.cu.
"""
from __future__ import annotations

from typing import List

import torch

from .arena import TILE, ParamArena, TensorTable
from .base import CommBackend, StepLog
from .window import DistBootstrap, Layout, SymmWindow, Window, wants_symm_window

ONE_SHOT_MAX_BYTES = 512 * 1024        # above this the all-reduce switches to two-shot


def default_timeout_ns(cfg=None) -> int:
    """Bound of every device-side peer wait (a wedged / dead peer trips the sticky status word instead of hanging
    the GPU): TrainConfig.peer_timeout_s, overridable with EGB_PEER_TIMEOUT_S."""
    import os
    sec = float(os.environ.get("EGB_PEER_TIMEOUT_S", getattr(cfg, "peer_timeout_s", 30.0) or 30.0))
    return int(sec * 1e9)


def _table_of(model) -> TensorTable:
    return TensorTable.from_named(list(model.named_parameters()), TILE)


def _wants_dbuf(cfg) -> bool:
    """csrc/gossip_dbuf.cu (two inbox slots, no WAR ack) applies to the dense, iter-sync, single-kernel decent step
    only; there it is the default (double_buffer=None), `--no-double-buffer` selects the single-slot + ack protocol."""
    db = getattr(cfg, "double_buffer", None)
    return ((db is None or bool(db)) and cfg.algo == "decent" and cfg.sync_mode == "iter"
            and not getattr(cfg, "overlap_push", False))


def build_layout(table: TensorTable, cfg, world: int, max_grid: int) -> Layout:
    lay = Layout()
    n = table.n_padded * 4
    lay.add("theta", n)
    lay.add("grad", n)
    if cfg.algo in ("decent", "event"):
        slots = 2 if _wants_dbuf(cfg) else 1      # double-buffered decent: slot = step & 1
        lay.add("inbox_l", n * slots)
        lay.add("inbox_r", n * slots)
    if cfg.algo == "spevent":
        K = sum(table.topk_counts(cfg.topk_percent))
        lay.add("rec_from_l", 2 * K * 4)
        lay.add("rec_from_r", 2 * K * 4)
        lay.add("seq_from_l", table.n_tensors * 4)
        lay.add("seq_from_r", table.n_tensors * 4)
        lay.add("done_from_l", 256)
        lay.add("done_from_r", 256)
    nflag = (table.n_tiles * 8 + 64) * 4      # one flag per (tile, warp)
    lay.add("flag_from_l", nflag)
    lay.add("flag_from_r", nflag)
    lay.add("ack_from_l", 256)
    lay.add("ack_from_r", 256)
    lay.add("pushed_from_l", 256)
    lay.add("pushed_from_r", 256)
    lay.add("ar_flags", 3 * max_grid * world * 4)
    return lay


def preallocate_arena_buffers(model, cfg, env, group=None, bootstrap=None):
    """Allocate this rank's window BEFORE the arena exists so that theta and grad live inside
    peer-mapped memory (cent reads peers' gradients; final averaging reads peers' theta)."""
    from ..ops import ext
    C = ext()
    table = _table_of(model)
    max_grid = C.gossip_max_grid(env.device.index or 0)
    if _wants_dbuf(cfg):
        max_grid = min(max_grid, C.gossip_dbuf_max_grid(env.device.index or 0))
    lay = build_layout(table, cfg, env.world, max_grid)
    symm = wants_symm_window(env) and bootstrap is None      # NVLS path: torch symmetric memory + multicast mapping
    win = (SymmWindow if symm else Window)(lay, env.rank, env.world, env.device)
    theta = win.view("theta", torch.float32)
    grad = win.view("grad", torch.float32)
    return theta, grad, {"window": win, "max_grid": max_grid, "bootstrap": bootstrap}


class P2PBackend(CommBackend):
    name = "p2p"
    zeroes_grad = True     # the fused kernels clear grad after consuming it

    def __init__(self, cfg, arena: ParamArena, ring, env, group=None, symm=None,
                 grid_cap: int = 0, defer_connect: bool = False, group_iters: int = 2,
                 timeout_ns: int = 0, vec256_push: bool = True, push_grid: int = 0):
        super().__init__(cfg, arena, ring)
        from ..ops import ext
        self.C = ext()
        if symm is None:
            raise ValueError("P2PBackend needs the window that holds the arena (preallocate_arena_buffers)")
        self.env, self.dev = env, arena.theta.device
        self.win: Window = symm["window"]
        self.boot = symm.get("bootstrap") or DistBootstrap(env, group)
        self.table = arena.table
        t = self.table
        if t.n_tensors > 1024:
            raise ValueError("p2p backend supports at most 1024 parameter tensors (EG_MAX_OWN)")
        self.grid = min(t.n_tiles, symm["max_grid"])
        if grid_cap:
            self.grid = min(self.grid, grid_cap)
        # balance: every CTA gets the same number of tiles (+-1 only on the last round)
        rounds = -(-t.n_tiles // self.grid)
        self.grid = -(-t.n_tiles // rounds)
        self.group_iters, self.timeout_ns, self.vec256_push = group_iters, timeout_ns or default_timeout_ns(cfg), vec256_push
        self.sync = cfg.sync_mode == "iter"
        self.gossip = cfg.algo in ("decent", "event", "spevent")
        self.sparse = cfg.algo == "spevent"
        self.do_comm = self.comm_enabled and self.gossip
        # split step: everything that depends only on theta_k (pushes / top-k records) is launched on a
        # side stream at the start of the step and overlaps forward+backward
        self.overlap = bool(getattr(cfg, "overlap_push", False)) and self.do_comm
        self.dbuf = _wants_dbuf(cfg) and self.do_comm
        self.ce_push = bool(getattr(cfg, "ce_push", False)) and self.overlap and cfg.algo == "decent" and self.sync
        self.nvls = self.nvls_step = False      # set in connect() when the window has a multicast mapping
        self.wire_dedup = False
        self.push_grid = push_grid
        self.recv_rms = cfg.dataset == "mnist"
        dev = self.dev
        i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=dev)
        zf = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        zi = lambda *s: torch.zeros(*s, dtype=torch.int32, device=dev)
        # ---- device tables ---------------------------------------------------------------
        self.d_tile_tensor = t.tile_to_tensor().to(dev)
        self.d_tile_start, self.d_tile_count, self.d_numel = i32(t.tile_start), i32(t.tile_count), i32(t.numels)
        if self.sparse:
            self.k = t.topk_counts(cfg.topk_percent)
            msg = [2 * k * 4 for k in self.k]
            rec_off, acc = [], 0
            for k in self.k:
                rec_off.append(acc)
                acc += 2 * k
            self.K = sum(self.k)
            self.d_k, self.d_rec_off = i32(self.k), i32(rec_off)
        else:
            msg = [n * 4 for n in t.numels]
        self.msg_bytes = msg
        self.d_msg = i32(msg)
        # ---- state -----------------------------------------------------------------------
        H = max(1, cfg.sent_history)
        self.thres, self.last_norm, self.last_iter = zf(t.n_tensors), zf(t.n_tensors), zf(t.n_tensors)
        self.slopes = zf(t.n_tensors * H)
        self.fire = torch.ones(t.n_tensors, dtype=torch.int32, device=dev)
        self.cur_norm = zf(t.n_tensors)
        self.counters = torch.zeros(4, dtype=torch.int64, device=dev)
        self.d_pass = zi(1)
        self.ticket = zi(8)      # [0] step tail, [1] push phase, [4] sparse, [6] all-reduce
        self.status = zi(1)
        self.tile_ss = zf(t.n_tiles * 8)
        self.tensor_done = zi(t.n_tensors)
        self.log_cap = 0
        self.log_ring = None
        self.tile_ss_l = self.tile_ss_r = None
        if self.want_logs and cfg.algo in ("event", "spevent"):
            self.log_cap = 4096
            self.log_ring = zf(self.log_cap * t.n_tensors * 5)
            self.tile_ss_l, self.tile_ss_r = zf(t.n_tiles * 8), zf(t.n_tiles * 8)
            self._drained = 0
            self._last_recv_l = torch.zeros(t.n_tensors)
            self._last_recv_r = torch.zeros(t.n_tensors)
        if self.sparse:
            # Q8: prev / replicas start from theta_0 (identical on all ranks)
            self.prev = arena.theta.clone()
            self.rep_l = arena.theta.clone()
            self.rep_r = arena.theta.clone()
            self.hist = torch.zeros(t.n_tensors * 2048, dtype=torch.int32, device=dev)
            self.sel_prefix, self.sel_remain = zi(t.n_tensors), zi(t.n_tensors)
            self.cand = torch.empty(t.n_padded, dtype=torch.int32, device=dev)      # candidate keys (worst case: all)
            self.cand_cnt, self.done1, self.done2 = zi(t.n_tensors), zi(t.n_tensors), zi(t.n_tensors)
            self.desc = torch.zeros(t.n_tiles, dtype=torch.int64, device=dev)       # look-back descriptors
            self.sp_bar = zi(4)
            self.applied_l, self.applied_r = zi(t.n_tensors), zi(t.n_tensors)
        self.ar_ctr = zi(1)
        # gradient pointer table (table mode): device arrays + rotating pinned staging
        self.table_mode = bool(getattr(arena, "table_mode", False)) and cfg.algo != "cent"
        self.d_gptr = torch.zeros(t.n_tensors, dtype=torch.int64, device=dev)
        self.d_gbf16 = zi(t.n_tensors)
        self._stage = []
        self._stage_i = 0
        self._grads_alive = None
        self.host_bytes = 0
        self._connected = False
        self.gp = self.ap = self.ap_avg = self.sp = None
        if hasattr(self.boot, "publish"):
            self.boot.publish("window_ptr", self.win.ptr)
            self.boot.publish("grid", self.grid)
        if not defer_connect:
            self.connect()

    # ------------------------------------------------------------------ wiring
    def connect(self) -> None:
        if self._connected:
            return
        boot, win, ring = self.boot, self.win, self.ring
        if hasattr(boot, "collect"):
            grids = boot.collect("grid")
        else:
            grids = boot.all_gather_object(self.grid)
        self.grid = int(min(grids))          # identical persistent grid on every rank
        boot.connect(win)
        L, R = ring.left, ring.right
        C, t, cfg = self.C, self.table, self.cfg
        a = self.arena
        P = lambda x: 0 if x is None else x.data_ptr()
        tab = {"tab.tile_tensor": P(self.d_tile_tensor), "tab.t_tile_start": P(self.d_tile_start),
               "tab.t_tile_count": P(self.d_tile_count), "tab.t_numel": P(self.d_numel),
               "tab.t_msg_bytes": P(self.d_msg), "tab.n_tiles": t.n_tiles, "tab.n_tensors": t.n_tensors}
        if True:   # the step kernel is also the plain fused-SGD path of cent / serial runs
            gp = C.GossipParams()
            dense = cfg.algo in ("decent", "event")
            gp.update(tab)
            gp.update({
                "theta": P(a.theta), "grad": P(a.grad), "mom": P(a.mom) if cfg.momentum != 0 else 0,
                "tile_ss": P(self.tile_ss), "tile_ss_l": P(self.tile_ss_l), "tile_ss_r": P(self.tile_ss_r),
                "shadow": P(getattr(a, "shadow", None)) if self.table_mode else 0,
                "t_grad_ptr": P(self.d_gptr) if self.table_mode else 0,
                "t_grad_bf16": P(self.d_gbf16) if self.table_mode else 0,
                "ticket": P(self.ticket), "tensor_done": P(self.tensor_done), "status": P(self.status),
                "timeout_ns": int(self.timeout_ns),
                "lr": float(cfg.lr), "mu": float(cfg.momentum),
                "need_norm": 1 if (cfg.algo in ("event", "spevent") or self.log_ring is not None) else 0,
                "do_mix": 1 if self.do_comm else 0,
                "do_push": 1 if (self.do_comm and dense) else 0,
                "sync": 1 if self.sync else 0,
                "send_ack": 1 if dense else 0,
                "zero_grad": 1, "group_iters": int(self.group_iters),
                "vec256_push": 1 if self.vec256_push else 0,
                "flag_from_l": win.addr("flag_from_l"), "flag_from_r": win.addr("flag_from_r"),
                "flag_to_l": win.addr("flag_from_r", L), "flag_to_r": win.addr("flag_from_l", R),
                "ack_from_l": win.addr("ack_from_l"), "ack_from_r": win.addr("ack_from_r"),
                "ack_to_l": win.addr("ack_from_r", L), "ack_to_r": win.addr("ack_from_l", R),
                "pushed_from_l": win.addr("pushed_from_l"), "pushed_from_r": win.addr("pushed_from_r"),
                "pushed_to_l": win.addr("pushed_from_r", L), "pushed_to_r": win.addr("pushed_from_l", R),
                "phase": 0,
                "fsm.thres": P(self.thres), "fsm.last_norm": P(self.last_norm),
                "fsm.last_iter": P(self.last_iter), "fsm.slopes": P(self.slopes),
                "fsm.fire": P(self.fire), "fsm.cur_norm": P(self.cur_norm),
                "fsm.counters": P(self.counters), "fsm.pass_num": P(self.d_pass),
                "fsm.log_ring": P(self.log_ring), "fsm.log_cap": int(self.log_cap),
                "fsm.horizon": float(cfg.horizon), "fsm.constant": float(cfg.constant),
                "fsm.thres_type": int(cfg.thres_type), "fsm.history": max(1, int(cfg.sent_history)),
                "fsm.initial_comm_passes": int(cfg.initial_comm_passes),
                "fsm.enabled": 1 if (cfg.algo in ("event", "spevent") and self.do_comm) else 0,
            })
            if dense:
                gp.update({"inbox_l": win.addr("inbox_l"), "inbox_r": win.addr("inbox_r"),
                           "push_l": win.addr("inbox_r", L), "push_r": win.addr("inbox_l", R)})
                # 2-rank ring: left == right.  Push theta ONCE (into the peer's inbox_r) and let the peer read that
                # copy as both L and R: (theta + 2 theta')/3 is unchanged bit for bit, the link carries 69.8 MB per
                # step instead of 139.6 MB.  Event/byte counters keep the reference's logical count (2 messages).
                self.wire_dedup = self.ring.world == 2 and L == R
                if self.wire_dedup:
                    gp.update({"inbox_l": win.addr("inbox_r"), "push_r": 0})
            elif self.sparse:
                gp.update({"inbox_l": P(self.rep_l), "inbox_r": P(self.rep_r)})
            self.gp = gp
        if self.sparse:
            sp = C.SparseParams()
            sp.update(tab)
            sp.update({
                "theta": P(a.theta), "prev": P(self.prev), "rep_l": P(self.rep_l), "rep_r": P(self.rep_r),
                "rec_from_l": win.addr("rec_from_l"), "rec_from_r": win.addr("rec_from_r"),
                "rec_to_l": win.addr("rec_from_r", L), "rec_to_r": win.addr("rec_from_l", R),
                "seq_from_l": win.addr("seq_from_l"), "seq_from_r": win.addr("seq_from_r"),
                "seq_to_l": win.addr("seq_from_r", L), "seq_to_r": win.addr("seq_from_l", R),
                "applied_l": P(self.applied_l), "applied_r": P(self.applied_r),
                "done_from_l": win.addr("done_from_l"), "done_from_r": win.addr("done_from_r"),
                "done_to_l": win.addr("done_from_r", L), "done_to_r": win.addr("done_from_l", R),
                "ack_from_l": win.addr("ack_from_l"), "ack_from_r": win.addr("ack_from_r"),
                "ack_to_l": win.addr("ack_from_r", L), "ack_to_r": win.addr("ack_from_l", R),
                "t_k": P(self.d_k), "t_rec_off": P(self.d_rec_off), "hist": P(self.hist),
                "sel_prefix": P(self.sel_prefix), "sel_remain": P(self.sel_remain),
                "cand": P(self.cand), "cand_cnt": P(self.cand_cnt), "done1": P(self.done1), "done2": P(self.done2),
                "desc": P(self.desc), "bar": P(self.sp_bar),
                "fire": P(self.fire), "pass_num": P(self.d_pass), "ticket": P(self.ticket[4:]),
                "status": P(self.status), "timeout_ns": int(self.timeout_ns), "sync": 1 if self.sync else 0,
            })
            # 2-rank ring: the one neighbour is both left and right -> write each record (and its seq / done flags)
            # ONCE, into the peer's from-right inbox, and let the peer read that copy for both replicas
            self.wire_dedup = self.ring.world == 2 and L == R
            if self.wire_dedup:
                sp.update({"rec_to_r": 0, "seq_to_r": 0, "done_to_r": 0,
                           "rec_from_l": win.addr("rec_from_r"), "seq_from_l": win.addr("seq_from_r"),
                           "done_from_l": win.addr("done_from_r")})
            self.sp = sp
            # the mix+SGD kernel runs the receive side (scatter into the replicas) as its prologue
            self.d_sp = torch.zeros((C.SPARSE_PARAMS_BYTES + 7) // 8, dtype=torch.int64, device=self.dev)
            C.sparse_params_to_device(sp, self.d_sp.data_ptr())
            if self.do_comm:
                self.gp.update({"sparse": self.d_sp.data_ptr()})
        # ---- all-reduce parameter blocks (cent step; final averaging for every algorithm) ----
        W = self.ring.world
        self.d_peer_grad = torch.tensor([win.addr("grad", r) for r in range(W)], dtype=torch.int64, device=self.dev)
        self.d_peer_theta = torch.tensor([win.addr("theta", r) for r in range(W)], dtype=torch.int64, device=self.dev)
        self.d_peer_flags = torch.tensor([win.addr("ar_flags", r) for r in range(W)], dtype=torch.int64, device=self.dev)
        common = {"peer_flags": P(self.d_peer_flags), "flags": win.addr("ar_flags"),
                  "ticket": P(self.ticket[6:]), "status": P(self.status), "step_ctr": P(self.ar_ctr),
                  "timeout_ns": int(self.timeout_ns), "n_tiles": t.n_tiles, "rank": self.ring.rank,
                  "world": W, "lr": float(cfg.lr), "mu": float(cfg.momentum)}
        ap = C.AllReduceParams()
        ap.update(common)
        ap.update({"peer_bufs": P(self.d_peer_grad), "local": P(a.grad), "theta": P(a.theta),
                   "mom": P(a.mom) if cfg.momentum != 0 else 0, "mode": 1, "zero_after": 1,
                   "two_shot": 1 if t.n_padded * 4 > ONE_SHOT_MAX_BYTES else 0})
        self.ap = ap
        av = C.AllReduceParams()
        av.update(common)
        av.update({"peer_bufs": P(self.d_peer_theta), "local": P(a.theta), "mode": 0, "zero_after": 0,
                   "two_shot": 1})
        self.ap_avg = av
        # EXPERIMENTAL NVLS: multicast mappings of the same buffers (0 unless the window is a SymmWindow on
        # a fabric with multicast support) -> csrc/allreduce_nvls.cu for the two-shot launches
        mc_grad = win.mc_addr("grad") if hasattr(win, "mc_addr") else 0
        mc_theta = win.mc_addr("theta") if hasattr(win, "mc_addr") else 0
        self.nvls = bool(mc_grad) and W > 1
        if self.nvls:
            ap.update({"mc_local": mc_grad})
            av.update({"mc_local": mc_theta})
        self.nvls_step = self.nvls and t.n_padded * 4 > ONE_SHOT_MAX_BYTES
        self._connected = True
        if self.gossip or self.table_mode:
            self._init_norms(run_fsm=self.gossip)
        boot.barrier()

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def _init_norms(self, run_fsm: bool) -> None:
        with torch.cuda.device(self.dev):
            self.C.gossip_init(self.gp, self.grid, 1 if run_fsm else 0, self._stream())

    # ------------------------------------------------------------------ gradient table
    def update_grad_table(self, compute, persistent: bool = False):
        """Point the step kernel at this backward's gradient tensors (one pointer + dtype flag per
        arena tensor).  Eager mode: every step, through rotating pinned staging buffers.  Graph
        capture: once -- the captured backward writes its gradients to fixed addresses; pass
        persistent=True and keep the returned staging buffers alive with the graph."""
        if not self.table_mode:
            return None
        n = self.table.n_tensors
        if persistent:
            hp, hf = torch.empty(n, dtype=torch.int64).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()
            ev = None
        else:
            if len(self._stage) < 4:
                self._stage.append((torch.empty(n, dtype=torch.int64).pin_memory(),
                                    torch.empty(n, dtype=torch.int32).pin_memory(), torch.cuda.Event()))
                hp, hf, ev = self._stage[-1]
            else:
                hp, hf, ev = self._stage[self._stage_i % 4]
                ev.synchronize()                  # the copy that last used this slot has executed
            self._stage_i += 1
        alive = []
        for i, c in enumerate(compute):
            g = c.grad
            if g is None:
                raise RuntimeError(f"tensor {self.table.names[i]} received no gradient")
            if g.stride() != c.stride() or g.dtype not in (torch.float32, torch.bfloat16):
                g = torch.empty_like(c, dtype=g.dtype if g.dtype in (torch.float32, torch.bfloat16)
                                     else torch.float32).copy_(g)
            hp[i] = g.data_ptr()
            hf[i] = 1 if g.dtype == torch.bfloat16 else 0
            alive.append(g)
        self.d_gptr.copy_(hp, non_blocking=True)
        self.d_gbf16.copy_(hf, non_blocking=True)
        if ev is not None:
            ev.record()
        self._grads_alive = alive                 # keep them until the next table replaces them
        return hp, hf, alive

    # ------------------------------------------------------------------ step
    graph_safe = True      # launch() only enqueues kernels whose per-step state lives on the device

    def launch_pre(self) -> None:
        """Split step, first half (side stream, concurrent with forward/backward): push theta_k to
        the neighbours / select + push the top-k records.  Depends only on theta_k and fire[]."""
        if not self.overlap:
            return
        C, s = self.C, self._stream()
        with torch.cuda.device(self.dev):
            if self.sparse:
                C.sparse_select_push(self.sp, self.grid, s)
            elif self.ce_push:
                C.ce_push(self.gp, s)                       # experimental: DMA engines instead of an SM kernel
            else:
                g = self.push_grid or max(1, min(self.grid, 128))
                C.gossip_step_phase(self.gp, 1, g, s)

    def launch(self) -> None:
        """Enqueue this step's kernels on the current stream (CUDA-graph capturable: the step
        counter, trigger state and handshake sequence numbers are all device resident)."""
        C, s = self.C, self._stream()
        with torch.cuda.device(self.dev):
            if self.cfg.algo == "cent":
                if self.ring.world > 1:
                    (C.allreduce_nvls if self.nvls_step else C.allreduce)(self.ap, self.grid, s)
                else:
                    C.gossip_step(self.gp, self.grid, s)       # plain fused SGD
                return
            if self.sparse and self.do_comm:
                if not self.overlap:
                    C.sparse_select_push(self.sp, self.grid, s)        # 3 launches: hist, candidates, compaction
                C.gossip_step(self.gp, self.grid, s)                   # receive prologue + mix + SGD + trigger
            elif self.overlap:
                C.gossip_step_phase(self.gp, 2, self.grid, s)
            elif self.dbuf:
                C.gossip_step_dbuf(self.gp, self.grid, s)
            else:
                C.gossip_step(self.gp, self.grid, s)

    def account_step(self) -> None:
        """Host-side bookkeeping of one executed step (after launch() or a graph replay)."""
        self.pass_num += 1
        if self.cfg.algo == "cent" and self.ring.world > 1:
            self.host_bytes += self.table.n_elems * 4
        elif self.cfg.algo == "decent" and self.do_comm:
            self.host_bytes += 2 * self.table.n_elems * 4

    def step(self) -> None:
        if self.overlap:        # no compute to hide behind when called stand-alone: run both halves in order
            self.launch_pre()
        self.launch()
        self.account_step()

    # ------------------------------------------------------------------ end of training
    def check_status(self) -> None:
        st = int(self.status.item())
        if st != 0:
            raise RuntimeError(f"p2p backend: device status {st} (1 = peer wait timed out) on rank {self.ring.rank}")

    def final_average_nocheck(self) -> None:
        if self.ring.world == 1:
            return
        with torch.cuda.device(self.dev):
            (self.C.allreduce_nvls if self.nvls else self.C.allreduce)(self.ap_avg, self.grid, self._stream())
        if not self.cfg.final_divide_all and self.ring.rank != 0:
            self.arena.theta.mul_(float(self.ring.world))     # reference quirk Q5: only rank 0 divides
        if self.table_mode and getattr(self.arena, "shadow", None) is not None:
            self._init_norms(run_fsm=False)                   # refresh the bf16 shadow of the averaged model

    def final_average(self) -> None:
        self.final_average_nocheck()
        self.check_status()

    def num_events(self) -> int:
        return int(self.counters[0].item())

    def total_events(self) -> int:
        ev = self.num_events()
        if self.ring.world == 1 or not hasattr(self.boot, "all_gather_object"):
            return ev
        return int(sum(self.boot.all_gather_object(ev)))

    def bytes_sent(self) -> int:
        return int(self.counters[1].item()) + self.host_bytes

    def synchronize(self) -> None:
        torch.cuda.synchronize(self.dev)

    # ------------------------------------------------------------------ logs
    def drain_logs(self) -> List[StepLog]:
        if self.log_ring is None:
            return []
        torch.cuda.synchronize(self.dev)
        t = self.table
        ring = self.log_ring.view(self.log_cap, t.n_tensors, 5).cpu()
        out = []
        numel = torch.tensor(t.numels, dtype=torch.float32)
        for s in range(self._drained + 1, self.pass_num + 1):
            row = ring[(s - 1) % self.log_cap]
            ln, rn = row[:, 3].clone(), row[:, 4].clone()
            if self.recv_rms:
                ln, rn = ln / numel.sqrt(), rn / numel.sqrt()
            lnew = (ln - self._last_recv_l).abs() > 0
            rnew = (rn - self._last_recv_r).abs() > 0
            self._last_recv_l = torch.where(lnew, ln, self._last_recv_l)
            self._last_recv_r = torch.where(rnew, rn, self._last_recv_r)
            out.append(StepLog(s, row[:, 0].clone(), row[:, 1].clone(), row[:, 2] > 0.5, ln, rn, lnew, rnew))
        self._drained = self.pass_num
        return out

    # ------------------------------------------------------------------ checkpoint
    def state_dict(self):
        torch.cuda.synchronize(self.dev)
        sd = {"pass_num": self.pass_num, "host_bytes": self.host_bytes}
        for k in ("thres", "last_norm", "last_iter", "slopes", "fire", "cur_norm", "counters", "d_pass"):
            sd[k] = getattr(self, k).cpu().clone()
        if self.cfg.algo in ("decent", "event"):
            sd["inbox_l"] = self.win.view("inbox_l", torch.float32).cpu().clone()
            sd["inbox_r"] = self.win.view("inbox_r", torch.float32).cpu().clone()
        if self.sparse:
            for k in ("prev", "rep_l", "rep_r"):
                sd[k] = getattr(self, k).cpu().clone()
        return sd

    def load_state_dict(self, sd):
        self.pass_num = int(sd["pass_num"])
        self.host_bytes = int(sd.get("host_bytes", 0))
        for k in ("thres", "last_norm", "last_iter", "slopes", "fire", "cur_norm", "counters", "d_pass"):
            getattr(self, k).copy_(sd[k].to(self.dev))
        if "inbox_l" in sd and self.cfg.algo in ("decent", "event"):
            self.win.view("inbox_l", torch.float32).copy_(sd["inbox_l"].to(self.dev))
            self.win.view("inbox_r", torch.float32).copy_(sd["inbox_r"].to(self.dev))
        if self.sparse:
            for k in ("prev", "rep_l", "rep_r"):
                getattr(self, k).copy_(sd[k].to(self.dev))
        if self.log_ring is not None:
            self._drained = self.pass_num          # log rows of the steps before the checkpoint are not replayed
        # NOTE: handshake counters (flags / acks) restart from pass_num on every rank
        if self.gossip:
            self._init_norms(run_fsm=False)
            self._reset_handshake(self.pass_num)
        torch.cuda.synchronize(self.dev)

    def _reset_handshake(self, step: int) -> None:
        for name in ("flag_from_l", "flag_from_r", "ack_from_l", "ack_from_r", "pushed_from_l", "pushed_from_r"):
            self.win.view(name, torch.int32).fill_(step)
        if self.sparse:
            for name in ("done_from_l", "done_from_r"):
                self.win.view(name, torch.int32).fill_(step)

    def close(self) -> None:
        try:
            torch.cuda.synchronize(self.dev)
        except Exception:
            pass
        self.win.close()
