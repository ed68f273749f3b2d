"""Dataset sources: synthetic MNIST/CIFAR-shaped data, and readers for the real files.

The reference reads MNIST through LibTorch's `datasets::MNIST(path)` + Normalize(0.1307,
0.3081) (/root/reference/dmnist/event/event.cpp:133-136) and CIFAR-10 from an image-folder
tree `train/<class>/<0000..4999>.jpg` decoded with OpenCV, resized to 32x32, BGR->RGB and
fed as raw 0..255 floats with no mean/std (/root/reference/dcifar10/common/custom.hpp:33-62,
:77-118, quirk Q7).  Here every source is materialised once as a uint8 [N,C,H,W] host tensor
(pinned when CUDA is present) + int64 labels; decoding, normalisation and augmentation
happen on the GPU after an async H2D copy of the raw bytes.

There is no network in the build environment, so `synthetic` is the default: class-
conditional templates + noise, which is learnable (tests assert the loss drops) and has the
exact shapes / dtypes / sizes of the real datasets.
"""
from __future__ import annotations

import gzip
import os
import pickle
import struct
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

CIFAR_CLASSES = ("airplane", "automobile", "bird", "cat", "deer", "dog", "frog", "horse",
                 "ship", "truck")  # custom.hpp:77-80 directory order


@dataclass
class DataSource:
    name: str
    images: torch.Tensor      # uint8 [N, C, H, W] on host
    labels: torch.Tensor      # int64 [N]
    scale: float              # x_float = x_u8 * scale
    mean: float
    std: float
    synthetic: bool = True

    def __len__(self) -> int:
        return int(self.images.shape[0])

    @property
    def sample_shape(self) -> Tuple[int, int, int]:
        return tuple(self.images.shape[1:])

    def pin(self) -> "DataSource":
        if torch.cuda.is_available():
            try:
                self.images = self.images.pin_memory()
                self.labels = self.labels.pin_memory()
            except RuntimeError:
                pass
        return self


def _norm_params(dataset: str):
    if dataset == "mnist":
        return 1.0 / 255.0, 0.1307, 0.3081          # ToTensor-style [0,1] then Normalize
    return 1.0, 0.0, 1.0                            # CIFAR: raw 0..255, un-normalised (Q7)


def synthetic_source(dataset: str, n: int, *, train: bool = True, seed: int = 1234,
                     classes: int = 10, noise: float = 48.0) -> DataSource:
    """Class-conditional synthetic images: one smooth random template per class + per-sample
    Gaussian noise, quantised to uint8. Deterministic in (dataset, n, train, seed)."""
    c, h, w = (1, 28, 28) if dataset == "mnist" else (3, 32, 32)
    g = torch.Generator().manual_seed(seed)                    # templates shared by train/test
    coarse = torch.rand(classes, c, 7, 7, generator=g)
    templates = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear",
                                                align_corners=False) * 255.0
    g2 = torch.Generator().manual_seed(seed + (1 if train else 2))
    labels = torch.randint(0, classes, (n,), generator=g2)
    images = torch.empty(n, c, h, w, dtype=torch.uint8)
    chunk = 8192
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = templates[labels[s:e]] + noise * torch.randn(e - s, c, h, w, generator=g2)
        images[s:e] = x.clamp_(0, 255).to(torch.uint8)
    sc, m, sd = _norm_params(dataset)
    return DataSource(f"synthetic-{dataset}", images, labels.long(), sc, m, sd, True)


# ----------------------------------------------------------------------------- real data
def _read_idx(path: str) -> np.ndarray:
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        magic, = struct.unpack(">I", f.read(4))
        nd = magic & 0xFF
        dims = struct.unpack(">" + "I" * nd, f.read(4 * nd))
        return np.frombuffer(f.read(), dtype=np.uint8).reshape(dims)


def _find(root: str, names) -> Optional[str]:
    for n in names:
        for cand in (n, n + ".gz"):
            p = os.path.join(root, cand)
            if os.path.exists(p):
                return p
    return None


def mnist_source(root: str, train: bool = True) -> DataSource:
    """MNIST idx files (the format LibTorch's datasets::MNIST reads)."""
    pre = "train" if train else "t10k"
    ip = _find(root, [f"{pre}-images-idx3-ubyte", f"{pre}-images.idx3-ubyte"])
    lp = _find(root, [f"{pre}-labels-idx1-ubyte", f"{pre}-labels.idx1-ubyte"])
    if ip is None or lp is None:
        raise FileNotFoundError(f"MNIST idx files not found under {root}")
    imgs = torch.from_numpy(_read_idx(ip).copy()).unsqueeze(1)
    labs = torch.from_numpy(_read_idx(lp).copy()).long()
    sc, m, sd = _norm_params("mnist")
    return DataSource("mnist", imgs, labs, sc, m, sd, False)


def cifar10_source(root: str, train: bool = True, shuffle_seed: int = 0) -> DataSource:
    """CIFAR-10 from (a) the reference's image-folder tree, (b) the python pickles or (c) the
    binary batches. The reference shuffles the file list once (custom.hpp:119-120); (a) mirrors
    that with a seeded permutation."""
    sc, m, sd = _norm_params("cifar10")
    split = "train" if train else "test"
    folder = os.path.join(root, split)
    if os.path.isdir(folder) and os.path.isdir(os.path.join(folder, CIFAR_CLASSES[0])):
        import cv2  # python OpenCV wheel
        imgs, labs = [], []
        for ci, cname in enumerate(CIFAR_CLASSES):
            d = os.path.join(folder, cname)
            for fn in sorted(os.listdir(d)):
                im = cv2.imread(os.path.join(d, fn), cv2.IMREAD_COLOR)
                if im is None:
                    continue
                im = cv2.cvtColor(cv2.resize(im, (32, 32)), cv2.COLOR_BGR2RGB)
                imgs.append(torch.from_numpy(im).permute(2, 0, 1))
                labs.append(ci)
        images, labels = torch.stack(imgs), torch.tensor(labs)
        perm = torch.randperm(len(labels), generator=torch.Generator().manual_seed(shuffle_seed))
        return DataSource("cifar10", images[perm].contiguous(), labels[perm].long(), sc, m, sd, False)
    pydir = root if os.path.exists(os.path.join(root, "data_batch_1")) else \
        os.path.join(root, "cifar-10-batches-py")
    if os.path.exists(os.path.join(pydir, "data_batch_1")):
        files = [f"data_batch_{i}" for i in range(1, 6)] if train else ["test_batch"]
        xs, ys = [], []
        for fn in files:
            with open(os.path.join(pydir, fn), "rb") as f:
                d = pickle.load(f, encoding="bytes")
            xs.append(np.asarray(d[b"data"], dtype=np.uint8).reshape(-1, 3, 32, 32))
            ys.extend(d[b"labels"])
        return DataSource("cifar10", torch.from_numpy(np.concatenate(xs)), torch.tensor(ys).long(),
                          sc, m, sd, False)
    bindir = root if os.path.exists(os.path.join(root, "data_batch_1.bin")) else \
        os.path.join(root, "cifar-10-batches-bin")
    if os.path.exists(os.path.join(bindir, "data_batch_1.bin")):
        files = [f"data_batch_{i}.bin" for i in range(1, 6)] if train else ["test_batch.bin"]
        raw = np.concatenate([np.fromfile(os.path.join(bindir, fn), dtype=np.uint8) for fn in files])
        raw = raw.reshape(-1, 3073)
        return DataSource("cifar10", torch.from_numpy(raw[:, 1:].reshape(-1, 3, 32, 32).copy()),
                          torch.from_numpy(raw[:, 0].astype(np.int64)), sc, m, sd, False)
    raise FileNotFoundError(f"no CIFAR-10 data (image folders / python / binary) under {root}")


def load_source(dataset: str, data: str, n: int, train: bool, seed: int = 1234) -> DataSource:
    if data == "synthetic":
        return synthetic_source(dataset, n, train=train, seed=seed).pin()
    src = mnist_source(data, train) if dataset == "mnist" else cifar10_source(data, train)
    if n and n < len(src):
        src.images, src.labels = src.images[:n].contiguous(), src.labels[:n].contiguous()
    return src.pin()
