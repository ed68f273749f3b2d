"""torch.distributed backend (gloo on CPU, NCCL on GPU): the BASELINE path.

It reproduces the reference algorithms with library collectives / point-to-point calls
plus eager elementwise ops -- i.e. what a straightforward port would look like:

  cent    all_reduce(grad)/R + SGD           (/root/reference/dmnist/cent/cent.cpp:130-145)
  decent  isend/irecv theta with both ring neighbours, (t+L+R)/3, SGD
                                            (/root/reference/dmnist/decent/decent.cpp:172-246)
  event   norm trigger -> send only fired tensors -> persistent inboxes -> mix -> SGD
                                            (/root/reference/dcifar10/event/event.cpp:282-479)
  spevent fired tensors travel as (top-k values, indices) records; receiver scatters into
          persistent neighbour replicas      (/root/reference/dcifar10/spevent/spevent.cpp:342-542)

torch.distributed has no one-sided RMA, so the "window" is emulated with an iteration-
synchronous exchange: first the fire masks, then a payload holding only the fired tensors.
Bytes on the wire therefore scale with events exactly as with MPI_Put, but every step is
a rendezvous (sync_mode=iter semantics).  This is the number the fused P2P backend beats;
on gloo it doubles as the GPU-free plumbing for BASELINE config 1.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist

from .base import CommBackend, StepLog
from .trigger import TriggerConfig, TriggerState, mix3_, sgd_, topk_select, trigger_step


class CollectiveBackend(CommBackend):
    name = "collective"

    def __init__(self, cfg, arena, ring, group=None):
        super().__init__(cfg, arena, ring)
        self.group = group
        self.name = dist.get_backend(group) if dist.is_initialized() else "local"
        t = arena.table
        dev = arena.theta.device
        self.tcfg = TriggerConfig.from_train(cfg)
        self.state = TriggerState(t.n_tensors, cfg.sent_history, dev)
        self.events = 0
        self.bytes = 0
        self.recv_rms = cfg.dataset == "mnist"      # MNIST logs RMS, CIFAR L2
        gossip = cfg.algo in ("decent", "event", "spevent")
        if gossip:
            if cfg.algo == "spevent":
                # Q8: replicas / prev start from theta_0 (identical on all ranks), not fresh random nets
                self.prev = arena.theta.clone()
                self.rep_l = arena.theta.clone()
                self.rep_r = arena.theta.clone()
                self.k = t.topk_counts(cfg.topk_percent)
            else:
                # RMA window halves, zero-initialised (event.cpp:144-147)
                self.inbox_l = torch.zeros_like(arena.theta)
                self.inbox_r = torch.zeros_like(arena.theta)
        if self.want_logs:
            z = lambda: torch.zeros(t.n_tensors, dtype=torch.float32, device=dev)
            self.last_recv_norm_l, self.last_recv_norm_r = z(), z()

    # ------------------------------------------------------------------ helpers
    def _flat(self, buf, i):
        return self.arena.flat(buf, i)

    def _exchange(self, to_left: List[torch.Tensor], to_right: List[torch.Tensor],
                  from_left: List[torch.Tensor], from_right: List[torch.Tensor]) -> None:
        """One rendezvous with both neighbours. When left == right (R == 2) the two payloads are
        identical by construction, so a single message is sent and used for both halves."""
        ring = self.ring
        if ring.world == 1:
            for d, s in zip(from_right, to_left):
                d.copy_(s)
            for d, s in zip(from_left, to_right):
                d.copy_(s)
            return
        ops = []
        if ring.left == ring.right:
            peer = ring.left
            for s in to_left:
                ops.append(dist.P2POp(dist.isend, s, peer, self.group))
            for d in from_left:
                ops.append(dist.P2POp(dist.irecv, d, peer, self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            for d, s in zip(from_right, from_left):
                d.copy_(s)
            return
        for s in to_left:
            ops.append(dist.P2POp(dist.isend, s, ring.left, self.group))
        for s in to_right:
            ops.append(dist.P2POp(dist.isend, s, ring.right, self.group))
        for d in from_left:
            ops.append(dist.P2POp(dist.irecv, d, ring.left, self.group))
        for d in from_right:
            ops.append(dist.P2POp(dist.irecv, d, ring.right, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _opt(self) -> None:
        a = self.arena
        sgd_(a.theta, a.grad, a.mom, self.cfg.lr, self.cfg.momentum)

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self) -> None:
        self.pass_num += 1
        algo = self.cfg.algo
        if algo == "cent":
            self._step_cent()
        elif not self.comm_enabled:
            self._opt()
        elif algo == "decent":
            self._step_decent()
        elif algo == "event":
            self._step_event()
        else:
            self._step_spevent()

    def _step_cent(self) -> None:
        a = self.arena
        if self.ring.world > 1:
            dist.all_reduce(a.grad, op=dist.ReduceOp.SUM, group=self.group)
            a.grad.div_(self.ring.world)
            self.bytes += a.table.n_elems * 4
        self._opt()

    def _step_decent(self) -> None:
        a = self.arena
        self._exchange([a.theta], [a.theta], [self.inbox_l], [self.inbox_r])
        self.bytes += 2 * a.table.n_elems * 4
        mix3_(a.theta, self.inbox_l, self.inbox_r)
        self._opt()

    def _fire(self):
        a = self.arena
        norms = a.tensor_norms()
        thres_before = None
        fire = trigger_step(self.state, norms, self.pass_num, self.tcfg)
        return norms, fire

    def _masks(self, fire: torch.Tensor):
        """Exchange fire masks with both neighbours; returns host bool lists (mine, left's, right's)."""
        m = fire.to(torch.uint8)
        ml, mr = torch.empty_like(m), torch.empty_like(m)
        self._exchange([m], [m], [ml], [mr])
        return fire.tolist(), ml.bool().tolist(), mr.bool().tolist()

    def _step_event(self) -> None:
        a, t = self.arena, self.arena.table
        thres_used = None
        if self.want_logs:
            # threshold as used in the comparison (after the horizon decay / constant set)
            tmp = self.state.clone()
        norms, fire = self._fire()
        mine, lm, rm = self._masks(fire)
        nfired = sum(mine)
        self.events += 2 * nfired
        sent = [self._flat(a.theta, i) for i in range(t.n_tensors) if mine[i]]
        payload = torch.cat(sent) if sent else a.theta.new_empty(0)
        self.bytes += 2 * payload.numel() * 4
        nl = sum(t.numels[i] for i in range(t.n_tensors) if lm[i])
        nr = sum(t.numels[i] for i in range(t.n_tensors) if rm[i])
        bl, br = a.theta.new_empty(nl), a.theta.new_empty(nr)
        self._exchange([payload], [payload], [bl], [br])
        for mask, buf, inbox in ((lm, bl, self.inbox_l), (rm, br, self.inbox_r)):
            off = 0
            for i in range(t.n_tensors):
                if mask[i]:
                    n = t.numels[i]
                    self._flat(inbox, i).copy_(buf[off:off + n])
                    off += n
        if self.want_logs:
            self._log(tmp, norms, fire, self.inbox_l, self.inbox_r)
        mix3_(a.theta, self.inbox_l, self.inbox_r)
        self._opt()

    def _step_spevent(self) -> None:
        a, t = self.arena, self.arena.table
        if self.want_logs:
            tmp = self.state.clone()
        norms, fire = self._fire()
        mine, lm, rm = self._masks(fire)
        self.events += 2 * sum(mine)
        vals, idxs = [], []
        for i in range(t.n_tensors):
            if mine[i]:
                th, pv = self._flat(a.theta, i), self._flat(self.prev, i)
                v, ix = topk_select(th, pv, self.k[i])
                pv[ix] = v                                   # spevent.cpp:407-413
                vals.append(v)
                idxs.append(ix.to(torch.int32))
        if vals:
            payload = torch.cat([torch.cat(vals), torch.cat(idxs).view(torch.float32)])
        else:
            payload = a.theta.new_empty(0)
        self.bytes += 2 * payload.numel() * 4
        kl = sum(self.k[i] for i in range(t.n_tensors) if lm[i])
        kr = sum(self.k[i] for i in range(t.n_tensors) if rm[i])
        bl, br = a.theta.new_empty(2 * kl), a.theta.new_empty(2 * kr)
        self._exchange([payload], [payload], [bl], [br])
        for mask, buf, ktot, rep in ((lm, bl, kl, self.rep_l), (rm, br, kr, self.rep_r)):
            off = 0
            for i in range(t.n_tensors):
                if mask[i]:
                    k = self.k[i]
                    v = buf[off:off + k]
                    ix = buf[ktot + off: ktot + off + k].view(torch.int32).long()
                    self._flat(rep, i)[ix] = v               # spevent.cpp:438-448
                    off += k
        if self.want_logs:
            self._log(tmp, norms, fire, self.rep_l, self.rep_r)
        mix3_(a.theta, self.rep_l, self.rep_r)
        self._opt()

    # ------------------------------------------------------------------ logging
    def _log(self, before: TriggerState, norms, fire, lbuf, rbuf) -> None:
        t = self.arena.table
        if self.tcfg.thres_type == 1:
            thres_used = before.thres * float(torch.tensor(self.tcfg.horizon, dtype=torch.float32))
        else:
            thres_used = torch.full_like(before.thres, self.tcfg.constant)
        ln = self.arena.tensor_norms(lbuf)
        rn = self.arena.tensor_norms(rbuf)
        if self.recv_rms:
            n = torch.tensor(t.numels, dtype=torch.float32, device=ln.device)
            ln, rn = ln / n.sqrt(), rn / n.sqrt()
        lnew = (ln - self.last_recv_norm_l).abs() > 0
        rnew = (rn - self.last_recv_norm_r).abs() > 0
        self.last_recv_norm_l = torch.where(lnew, ln, self.last_recv_norm_l)
        self.last_recv_norm_r = torch.where(rnew, rn, self.last_recv_norm_r)
        self.logs.append(StepLog(self.pass_num, norms.cpu(), thres_used.cpu(), fire.cpu(),
                                 ln.cpu(), rn.cpu(), lnew.cpu(), rnew.cpu()))

    # ------------------------------------------------------------------ end of training
    @torch.no_grad()
    def final_average(self) -> None:
        """All-reduce the parameters (event.cpp:509-519). The reference divides on rank 0 only
        (Q5); by default every rank divides so all ranks hold the averaged model."""
        if self.ring.world == 1:
            return
        a = self.arena
        dist.all_reduce(a.theta, op=dist.ReduceOp.SUM, group=self.group)
        if self.cfg.final_divide_all or self.ring.rank == 0:
            a.theta.div_(self.ring.world)

    def num_events(self) -> int:
        return int(self.events)

    def total_events(self) -> int:
        if self.ring.world == 1:
            return int(self.events)
        e = torch.tensor([self.events], dtype=torch.int64, device=self.arena.theta.device)
        dist.all_reduce(e, group=self.group)
        return int(e.item())

    def bytes_sent(self) -> int:
        return int(self.bytes)

    def synchronize(self) -> None:
        if self.arena.theta.is_cuda:
            torch.cuda.synchronize()

    def state_dict(self):
        sd = {"pass_num": self.pass_num, "events": self.events, "bytes": self.bytes,
              "fsm": self.state.state_dict()}
        for k in ("inbox_l", "inbox_r", "prev", "rep_l", "rep_r"):
            if hasattr(self, k):
                sd[k] = getattr(self, k).cpu().clone()
        return sd

    def load_state_dict(self, sd):
        self.pass_num = int(sd["pass_num"])
        self.events = int(sd["events"])
        self.bytes = int(sd.get("bytes", 0))
        self.state.load_state_dict(sd["fsm"])
        for k in ("inbox_l", "inbox_r", "prev", "rep_l", "rep_r"):
            if hasattr(self, k) and k in sd:
                getattr(self, k).copy_(sd[k].to(getattr(self, k).device))
