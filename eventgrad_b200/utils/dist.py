"""Process-group bootstrap: one process per GPU (torchrun), NCCL on GPU / gloo on CPU.

Replaces `mpirun -np N` + MPI_Init/Comm_rank/Comm_size
(/root/reference/dmnist/event/event.cpp:108-110).  The process group is only plumbing
(rendezvous, IPC-handle exchange, final statistics, the NCCL baseline backend); the hot
path of the p2p backend never calls it.
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class DistEnv:
    rank: int
    world: int
    local_rank: int
    device: torch.device
    backend: str          # nccl | gloo | none


def init_distributed(device_pref: str = "auto", timeout_s: int = 0) -> DistEnv:
    """timeout_s bounds every collective of the process group (0: $EGB_DIST_TIMEOUT or 600 s): a dead
    rank surfaces as an exception on its peers instead of the reference's indefinite MPI_Recv hang."""
    if timeout_s <= 0:
        timeout_s = int(os.environ.get("EGB_DIST_TIMEOUT", "600"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    use_cuda = (device_pref == "cuda") or (device_pref == "auto" and torch.cuda.is_available())
    if use_cuda:
        device = torch.device("cuda", local_rank % max(1, torch.cuda.device_count()))
        torch.cuda.set_device(device)
    else:
        device = torch.device("cpu")
    backend = "none"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        backend = "nccl" if use_cuda else "gloo"
        if not dist.is_initialized():
            kw = {}
            if use_cuda:
                kw["device_id"] = device
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    timeout=datetime.timedelta(seconds=timeout_s), **kw)
    return DistEnv(rank, world, local_rank, device, backend)


def barrier(env: DistEnv) -> None:
    if env.world > 1:
        if env.device.type == "cuda":
            dist.barrier(device_ids=[env.device.index])
        else:
            dist.barrier()


def max_over_ranks(value: float, env: DistEnv) -> float:
    if env.world == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=env.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, env: DistEnv) -> float:
    if env.world == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=env.device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def shutdown() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()
