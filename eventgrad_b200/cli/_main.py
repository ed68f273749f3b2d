"""Shared main() of the five entrypoints: parse -> init process group -> train -> finalize."""
from __future__ import annotations

import json
from typing import Optional, Sequence

from ..config import parse_cli
from ..engine.trainer import Trainer
from ..utils.dist import init_distributed, shutdown


def run(program: str, argv: Optional[Sequence[str]] = None) -> dict:
    cfg = parse_cli(program, argv)
    env = init_distributed(cfg.device)
    tr = Trainer(cfg, env)
    try:
        tr.fit()
        res = tr.finalize()
    finally:
        tr.close()
    if env.rank == 0 and not cfg.quiet:
        keep = {k: res[k] for k in ("events_total", "dense_messages", "messages_saved",
                                    "train_time_s", "steps") if k in res}
        print("summary " + json.dumps(keep), flush=True)
    shutdown()
    return res
