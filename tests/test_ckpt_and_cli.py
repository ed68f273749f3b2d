import os

import torch

from eventgrad_b200.config import parse_cli, preset
from eventgrad_b200.data import synthetic_source
from eventgrad_b200.engine.trainer import Trainer
from eventgrad_b200.parallel.topology import Ring
from eventgrad_b200.utils.ckpt import ckpt_path
from eventgrad_b200.utils.dist import DistEnv


def _env():
    return DistEnv(0, 1, 0, torch.device("cpu"), "none")


def test_positional_cli_compat():
    c = parse_cli("cifar_spevent", ["1", "1", "0.9", "5"])
    assert (c.file_write, c.thres_type, c.horizon, c.topk_percent, c.algo) == (1, 1, 0.9, 5.0, "spevent")
    c = parse_cli("mnist_event", ["0", "0", "0.01"])
    assert c.thres_type == 0 and c.constant == 0.01 and c.model == "cnn2" and c.batch_size == 64
    c = parse_cli("cent", [])
    assert c.epochs == 250 and c.batch_mode == "full" and c.lr == 1e-2
    c = parse_cli("decent", ["1", "--epochs", "3"])
    assert c.file_write == 1 and c.epochs == 3 and c.sampler == "sequential"
    c = parse_cli("cifar_event", [])
    assert (c.batch_size, c.momentum, c.epochs, c.model, c.resnet_variant) == (256, 0.9, 20, "resnet18", "ref")


def test_ring_edge_cases():
    assert Ring(0, 1).neighbours() == (0, 0) and Ring(0, 1).serial
    assert Ring(0, 2).neighbours() == (1, 1) and Ring(0, 2).degenerate_pair
    assert Ring(0, 8).neighbours() == (7, 1) and Ring(7, 8).neighbours() == (6, 0)
    assert abs(sum(Ring(3, 8).mixing_row()) - 1.0) < 1e-12


def test_serial_training_learns_and_checkpoint_resumes_exactly(tmp_path):
    cfg = preset("mnist_event", device="cpu", backend="gloo", train_samples=1024, test_samples=256,
                 epochs=2, quiet=True, ckpt_dir=str(tmp_path), ckpt_every=1)
    src = synthetic_source("mnist", 1024)
    torch.manual_seed(0)
    tr = Trainer(cfg, _env(), train_source=src, test_source=synthetic_source("mnist", 256, train=False))
    tr.fit()
    res = tr.finalize()
    assert res["test_acc"] > 30.0                                   # synthetic classes are learnable
    assert res["events_total"] > 0 and res["events_total"] % 2 == 0
    path = ckpt_path(str(tmp_path), 0)
    assert os.path.exists(path)
    # resume: state after epoch 2 must be restored bit-for-bit
    cfg2 = cfg.replace(resume=path, epochs=2)
    tr2 = Trainer(cfg2, _env(), train_source=src)
    assert tr2.epoch == 2 and tr2.backend.pass_num == tr.backend.pass_num
    assert torch.equal(tr2.arena.theta, torch.load(path, weights_only=False)["arena"]["theta"])
    assert torch.equal(tr2.backend.state.thres, tr.backend.state.thres)
