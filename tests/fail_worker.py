"""Failure-detection worker: rank 1 dies after a few steps; the survivors must raise, not hang."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventgrad_b200.config import TrainConfig  # noqa: E402
from eventgrad_b200.models import build_model  # noqa: E402
from eventgrad_b200.parallel import ParamArena, Ring, make_backend  # noqa: E402
from eventgrad_b200.utils.dist import init_distributed  # noqa: E402

env = init_distributed("cpu")
cfg = TrainConfig(algo="decent", dataset="mnist", model="cnn2", backend="gloo").validate()
torch.manual_seed(0)
arena = ParamArena(build_model("cnn2"), env.device)
be = make_backend(cfg, arena, Ring(env.rank, env.world), env)
t0 = time.time()
try:
    for s in range(50):
        if env.rank == 1 and s == 3:
            os._exit(17)                       # simulated crash: no clean shutdown
        arena.grad.normal_(0, 0.01)
        be.step()
    print("UNEXPECTED_COMPLETION", flush=True)
    sys.exit(3)
except Exception as e:  # noqa: BLE001
    print(f"PEER_FAILURE_DETECTED rank={env.rank} after {time.time() - t0:.1f}s: {type(e).__name__}", flush=True)
    os._exit(0)
