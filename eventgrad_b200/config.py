"""Configuration: one dataclass + argparse, with a positional-argv compatibility
shim for the reference programs' CLI.

Reference CLI (positional, no validation):
  dmnist/event, dcifar10/event : <file_write> <thres_type> <horizon|constant>
      (/root/reference/dmnist/event/event.cpp:89-100,
       /root/reference/dcifar10/event/event.cpp:46-57)
  dcifar10/spevent             : ... + <topk_percent>
      (/root/reference/dcifar10/spevent/spevent.cpp:47-60)
  dmnist/decent                : <file_write>   (decent.cpp:44)
  dmnist/cent                  : none
Every other knob in the reference is a source constant;
here they are the *defaults* of the matching preset and can be overridden.
"""
from __future__ import annotations

import argparse
import dataclasses
from dataclasses import dataclass
from typing import Optional, Sequence

ALGOS = ("cent", "decent", "event", "spevent")
DATASETS = ("mnist", "cifar10")


@dataclass
class TrainConfig:
    # ---- what to run -----------------------------------------------------
    algo: str = "event"              # cent | decent | event | spevent
    dataset: str = "cifar10"         # mnist | cifar10
    model: str = "resnet18"          # mlp | cnn1 | cnn2 | lenet | resnet18/34/50/101/152
    resnet_variant: str = "ref"      # ref (blocks+1 per stage, as shipped) | canonical
    # ---- optimisation ----------------------------------------------------
    epochs: int = 20
    batch_size: int = 256            # meaning depends on batch_mode
    batch_mode: str = "global"       # global (split over ranks) | per_rank | full (whole shard)
    lr: float = 1e-2
    momentum: float = 0.9
    seed: int = 0
    # ---- event trigger -------------------------------------
    thres_type: int = 1              # 1 adaptive (horizon), 0 constant
    horizon: float = 1.0
    constant: float = 0.0
    sent_history: int = 2
    initial_comm_passes: int = 30
    topk_percent: float = 10.0       # spevent only
    # ---- communication ---------------------------------------------------
    backend: str = "auto"            # auto | p2p (fused sm_100a kernels) | nccl | gloo | refport (reference-structured)
    sync_mode: str = "iter"          # iter (deterministic handshake) | async (reference RMA semantics)
    final_divide_all: bool = True    # reference divides on rank 0 only (Q5)
    grad_table: bool = True          # p2p gossip: step kernel reads autograd's gradients in place (+ bf16
                                     # shadow weights under dtype=bf16) instead of an fp32 grad arena
    overlap_push: Optional[bool] = None   # p2p: launch the push half of the step on a side stream so it overlaps
                                     # forward/backward (False = single fused kernel).  None = auto: on for
                                     # decent/event on >= 2 GPUs (measured faster at N=2/4/8, profiles/README.md)
    spevent_fresh_replicas: bool = False   # reference quirk Q8: prev/left/right replicas of spevent are three
                                     # MORE randomly initialised networks (spevent.cpp:123-136) instead of theta_0
    ce_push: bool = False            # decent + overlap_push: the push half of the split step is copy-engine memcpys
                                     # instead of an SM kernel (csrc/ce_push.cu)
    double_buffer: Optional[bool] = None   # decent, p2p, iter-sync, fused (non-split) step: two inbox slots and no
                                     # WAR ack (csrc/gossip_dbuf.cu).  None = on where it applies
    peer_timeout_s: float = 30.0     # bound of every device-side peer wait (sticky status instead of a hang);
                                     # EGB_PEER_TIMEOUT_S overrides
    # ---- data --------------------------------------------------------------
    data: str = "synthetic"          # synthetic | path to dataset root
    sampler: str = "random"          # random | sequential
    augment: bool = True             # pad4 + flip + crop (cifar only)
    native_loader: str = "auto"      # auto (CPU: native C++ prefetcher, CUDA: Python staging) | on | off
    train_samples: int = 50000
    test_samples: int = 10000
    test_batch_size: int = 100
    # ---- execution ---------------------------------------------------------
    device: str = "auto"             # auto | cuda | cpu
    dtype: str = "fp32"              # fp32 | tf32 | bf16 (compute dtype; arena is fp32)
    channels_last: Optional[bool] = None   # None = auto on CUDA: NHWC, except fp32 with conv_tc off (cuDNN's fp32 convs are NCHW)
    conv_tc: Optional[bool] = None         # fp32 only: 3x3 convs on the tcgen05 tensor cores at fp32 accuracy (csrc/conv_tc.cu);
                                           # None = on for CUDA (EGB_CONV_TC=0 turns it off); False = cuDNN's SIMT fp32 kernels
    cuda_graph: Optional[bool] = None      # None = auto: whole-step CUDA graph on CUDA
    max_steps: int = 0               # >0: stop after this many steps (tests / bench)
    cudnn_benchmark: bool = True     # cuDNN autotune: best steady state, but every new conv shape costs a
                                     # one-time search (seconds; it also hits the partial last batch)
    host_threads: int = 2            # intra-op CPU threads on GPU runs (0 = leave torch's default)
    # ---- observability -------------------------------------------------------
    file_write: int = 0              # reference debug files send/recv/train/values<r>.txt
    log_dir: str = "."
    ckpt_dir: str = ""
    ckpt_every: int = 0              # epochs; 0 = never
    resume: str = ""
    quiet: bool = False
    phase_timers: bool = False       # per-phase device timers (CUDA events + NVTX ranges), printed by rank 0

    def validate(self) -> "TrainConfig":
        if self.algo not in ALGOS:
            raise ValueError(f"algo must be one of {ALGOS}, got {self.algo!r}")
        if self.dataset not in DATASETS:
            raise ValueError(f"dataset must be one of {DATASETS}")
        if self.sync_mode not in ("iter", "async"):
            raise ValueError("sync_mode must be iter|async")
        if self.algo == "decent" and self.sync_mode != "iter":
            raise ValueError("decent (D-PSGD) is lock-step by definition: sync_mode=iter")
        if self.batch_mode not in ("global", "per_rank", "full"):
            raise ValueError("batch_mode must be global|per_rank|full")
        if self.thres_type not in (0, 1):
            raise ValueError("thres_type must be 0 (constant) or 1 (adaptive)")
        if not (0.0 < self.topk_percent <= 100.0):
            raise ValueError("topk_percent must be in (0, 100]")
        if self.dtype not in ("fp32", "tf32", "bf16"):
            raise ValueError("dtype must be fp32|tf32|bf16")
        return self

    def resolved(self, device_type: str, world: int) -> "TrainConfig":
        """Fill the auto (None) execution switches for the device the job actually runs on: on a GPU the fast path
        (NHWC + fused BN kernels, whole-step CUDA graph, pushes overlapped with backward) IS the default path."""
        cuda = device_type == "cuda"
        p2p = self.backend == "p2p" or (self.backend == "auto" and cuda)
        kw = {}
        conv_tc = self.conv_tc
        if conv_tc is None:
            import os
            conv_tc = kw["conv_tc"] = bool(cuda and self.dtype == "fp32" and os.environ.get("EGB_CONV_TC", "1") != "0")
        if self.channels_last is None:
            # bf16 / tf32 / fp32 on our tensor-core convolutions: NHWC.  fp32 on cuDNN (conv_tc off): its IEEE-fp32
            # convolutions are NCHW kernels -- fed NHWC they transpose around every conv (measured 36.9 vs 27.8 ms/step
            # on B200) -- so that path stays NCHW; the fused BN kernels exist for both layouts
            kw["channels_last"] = cuda and (self.dtype != "fp32" or conv_tc)
        if self.cuda_graph is None:
            kw["cuda_graph"] = cuda
        if self.overlap_push is None:
            kw["overlap_push"] = bool(cuda and p2p and world >= 2 and self.algo in ("decent", "event")
                                      and not bool(self.double_buffer))
        return dataclasses.replace(self, **kw) if kw else self

    def replace(self, **kw) -> "TrainConfig":
        return dataclasses.replace(self, **kw)


# Presets = the constants hard-coded in each reference program.
PRESETS = {
    # dmnist/cent/cent.cpp:62-65,75,95  (full batch per rank, lr 1e-2, 250 epochs, random sampler)
    "cent": dict(algo="cent", dataset="mnist", model="mlp", epochs=250, batch_mode="full",
                 lr=1e-2, momentum=0.0, sampler="random", train_samples=60000, augment=False),
    # dmnist/decent/decent.cpp:83-86,121,139 (full batch, lr 1e-2, 50 epochs, sequential)
    "decent": dict(algo="decent", dataset="mnist", model="mlp", epochs=50, batch_mode="full",
                   lr=1e-2, momentum=0.0, sampler="sequential", train_samples=60000,
                   augment=False, sync_mode="iter"),
    # dmnist/event/event.cpp:145,227-230,255 (batch 64 per rank, lr 0.05, 10 epochs, sequential)
    "mnist_event": dict(algo="event", dataset="mnist", model="cnn2", epochs=10, batch_size=64,
                        batch_mode="per_rank", lr=0.05, momentum=0.0, sampler="sequential",
                        train_samples=60000, augment=False),
    # dcifar10/event/event.cpp:29-42,91,196-200
    "cifar_event": dict(algo="event", dataset="cifar10", model="resnet18", epochs=20,
                        batch_size=256, batch_mode="global", lr=1e-2, momentum=0.9,
                        sampler="random", train_samples=50000, augment=True),
    # dcifar10/spevent/spevent.cpp:32-37,60,219-223
    "cifar_spevent": dict(algo="spevent", dataset="cifar10", model="resnet18", epochs=20,
                          batch_size=256, batch_mode="global", lr=1e-2, momentum=0.9,
                          sampler="random", train_samples=50000, augment=True),
}


def preset(name: str, **overrides) -> TrainConfig:
    base = dict(PRESETS[name])
    base.update(overrides)
    return TrainConfig(**base).validate()


def _add_common_flags(p: argparse.ArgumentParser) -> None:
    d = TrainConfig()
    p.add_argument("--model", default=None)
    p.add_argument("--resnet-variant", default=None, choices=["ref", "canonical"])
    p.add_argument("--epochs", type=int, default=None)
    p.add_argument("--batch-size", type=int, default=None)
    p.add_argument("--batch-mode", default=None, choices=["global", "per_rank", "full"])
    p.add_argument("--lr", type=float, default=None)
    p.add_argument("--momentum", type=float, default=None)
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--sent-history", type=int, default=None)
    p.add_argument("--initial-comm-passes", type=int, default=None)
    p.add_argument("--backend", default=None, choices=["auto", "p2p", "nccl", "gloo", "refport"])
    p.add_argument("--sync-mode", default=None, choices=["iter", "async"])
    p.add_argument("--overlap-push", action=argparse.BooleanOptionalAction, default=None,
                   help="default: on for decent/event on >= 2 GPUs")
    p.add_argument("--double-buffer", action=argparse.BooleanOptionalAction, default=None)
    p.add_argument("--peer-timeout-s", type=float, default=None)
    p.add_argument("--ce-push", action="store_true", default=None)
    p.add_argument("--fresh-replicas", dest="spevent_fresh_replicas", action="store_true", default=None,
                   help="spevent: initialise prev/left/right replicas like the reference (three more random nets)")
    p.add_argument("--no-grad-table", dest="grad_table", action="store_false", default=None)
    p.add_argument("--data", default=None, help="'synthetic' or dataset root directory")
    p.add_argument("--sampler", default=None, choices=["random", "sequential"])
    p.add_argument("--no-augment", dest="augment", action="store_false", default=None)
    p.add_argument("--native-loader", default=None, choices=["auto", "on", "off"])
    p.add_argument("--train-samples", type=int, default=None)
    p.add_argument("--test-samples", type=int, default=None)
    p.add_argument("--device", default=None, choices=["auto", "cuda", "cpu"])
    p.add_argument("--dtype", default=None, choices=["fp32", "tf32", "bf16"])
    p.add_argument("--channels-last", action=argparse.BooleanOptionalAction, default=None, help="default: on for CUDA")
    p.add_argument("--cuda-graph", action=argparse.BooleanOptionalAction, default=None, help="default: on for CUDA")
    p.add_argument("--conv-tc", action=argparse.BooleanOptionalAction, default=None,
                   help="fp32: convolutions on the tcgen05 tensor cores at fp32 accuracy (default: on for CUDA); "
                        "--no-conv-tc = cuDNN's SIMT fp32 kernels")
    p.add_argument("--no-cudnn-benchmark", dest="cudnn_benchmark", action="store_false", default=None)
    p.add_argument("--max-steps", type=int, default=None)
    p.add_argument("--log-dir", default=None)
    p.add_argument("--ckpt-dir", default=None)
    p.add_argument("--ckpt-every", type=int, default=None)
    p.add_argument("--resume", default=None)
    p.add_argument("--quiet", action="store_true", default=None)
    p.add_argument("--phase-timers", action="store_true", default=None)
    del d


def parse_cli(program: str, argv: Optional[Sequence[str]] = None) -> TrainConfig:
    """Parse the CLI of one of the five reference-equivalent programs.

    `program` is one of cent | decent | mnist_event | cifar_event | cifar_spevent.
    Positional arguments follow the reference order; everything else is a flag.
    """
    p = argparse.ArgumentParser(prog=f"eventgrad_b200.cli.{program}",
                                description=f"{program}: reference-compatible positional args + flags")
    if program in ("mnist_event", "cifar_event", "cifar_spevent"):
        p.add_argument("file_write", type=int, nargs="?", default=0)
        p.add_argument("thres_type", type=int, nargs="?", default=1,
                       help="0 constant threshold, 1 adaptive")
        p.add_argument("threshold_arg", type=float, nargs="?", default=1.0,
                       help="horizon if thres_type==1 else the constant threshold")
        if program == "cifar_spevent":
            p.add_argument("topk_percent", type=float, nargs="?", default=10.0)
    elif program == "decent":
        p.add_argument("file_write", type=int, nargs="?", default=0)
    _add_common_flags(p)
    ns = p.parse_args(argv)
    over = {}
    for f in dataclasses.fields(TrainConfig):
        v = getattr(ns, f.name, None)
        if v is not None:
            over[f.name] = v
    if hasattr(ns, "threshold_arg"):
        if over.get("thres_type", 1) == 1:
            over["horizon"] = ns.threshold_arg
        else:
            over["constant"] = ns.threshold_arg
    return preset(program, **over)
