#!/usr/bin/env python
"""Run a drop-in Runner on synthetic data with the reference's CLI surface (demo_vision.py:16-56): the same
``--method`` / ``--hparams`` / ``--backbone`` / ``--lr`` / ``--lr_head`` / ``--momentum`` / ``--epochs`` flags, but the
loaders are synthetic tensors of the BASELINE.json shapes (no dataset download) and the backbones are random-init.

    python examples/demo_synthetic.py --backbone mlp_mnist --method sgld \
        --hparams prior_sig=1.0,Ninflate=1e3,nd=1.0,burnin=1,thin=10,bias=informative,nst=5 --lr 1e-2 --momentum 0.5
    python examples/demo_synthetic.py --backbone resnet101 --method csghmc --num_cycles 2 --epochs 4 \
        --hparams prior_sig=1.0,Ninflate=1.0,nd=0.01,burnin=0,momentum_decay=0.18,thin=2,bias=informative,nst=2 \
        --lr 1e-4 --lr_head 1e-2 --batch_size 16 --train_batches 8 --pretrained synthetic
    python examples/demo_synthetic.py --backbone resnet101 --method csghmc_fs --num_cycles 2 --epochs 6 \
        --hparams prior_sig=1.0,Ninflate=1.0,nd=0.01,burnin=0,momentum_decay=0.18,thin=2,bias=informative,nst=2 \
        --lr 1e-4 --lr_head 1e-2 --batch_size 16 --train_batches 6     # raw samples in the HBM ring + BMA
"""
import argparse
import importlib
import logging
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesdll_b200 import shapes  # noqa: E402

METHODS = ("sgld", "sghmc", "csgld", "csghmc", "csghmc_fs", "adam_sghmc", "adam_csghmc")


def synthetic_loader(n_batches, batch, shape, num_classes, seed):
    gen = torch.Generator().manual_seed(seed)
    return [(torch.randn(batch, *shape, generator=gen).pin_memory(),
             torch.randint(0, num_classes, (batch,), generator=gen).pin_memory()) for _ in range(n_batches)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--method", default="sghmc", choices=METHODS)
    ap.add_argument("--hparams", default="prior_sig=1.0,Ninflate=1e3,nd=1.0,momentum_decay=0.18,burnin=1,thin=1,bias=informative,nst=5")
    ap.add_argument("--pretrained", default=None, help="any non-None string: use a (random-init) net0 as the prior mean")
    ap.add_argument("--backbone", default="mlp_mnist", choices=("mlp_mnist", "resnet101", "vit_l_32"))
    ap.add_argument("--ece_num_bins", type=int, default=15)
    ap.add_argument("--num_cycles", type=int, default=1)
    ap.add_argument("--proportion_exploration", type=float, default=0.5)
    ap.add_argument("--full_sample", action="store_true")
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--batch_size", type=int, default=128)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--lr_head", type=float, default=None)
    ap.add_argument("--momentum", type=float, default=0.5)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--log_dir", default=None)
    ap.add_argument("--test_eval_freq", type=int, default=1)
    ap.add_argument("--train_batches", type=int, default=10)
    ap.add_argument("--test_batches", type=int, default=4)
    args = ap.parse_args()

    args.device = torch.device("cuda")
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    args.hparams = dict(kv.split("=") for kv in args.hparams.replace('"', "").split(",") if "=" in kv)
    if args.lr_head is None:
        args.lr_head = args.lr
    args.log_dir = args.log_dir or tempfile.mkdtemp(prefix=f"bdl_{args.method}_")
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s", stream=sys.stderr)
    logger = logging.getLogger("demo")

    mnist = args.backbone == "mlp_mnist"
    args.num_classes = 10 if mnist else 37
    shape = (1, 28, 28) if mnist else (3, 224, 224)
    train = synthetic_loader(args.train_batches, args.batch_size, shape, args.num_classes, args.seed)
    val = synthetic_loader(args.test_batches, args.batch_size, shape, args.num_classes, args.seed + 1)
    test = synthetic_loader(args.test_batches, args.batch_size, shape, args.num_classes, args.seed + 2)
    args.ND = args.train_batches * args.batch_size

    net = shapes.create_backbone(args.backbone, args.num_classes)
    net0 = shapes.create_backbone(args.backbone, args.num_classes) if args.pretrained is not None else None
    Runner = importlib.import_module(f"bayesdll_b200.methods.{args.method}").Runner
    runner = Runner(net, net0, args, logger)
    t0 = time.time()
    out = runner.train(train, val, test)
    torch.cuda.synchronize()
    n = sum(p.numel() for p in runner.net.parameters())
    logger.info(f"done: {args.method} on {args.backbone} ({n} params), {args.epochs} epochs x {args.train_batches} batches in "
                f"{time.time() - t0:.1f} s; outputs in {args.log_dir}: {sorted(os.listdir(args.log_dir))}")
    if isinstance(out, dict):
        logger.info(f"samples per cycle: {out['samples_per_cycle']}")


if __name__ == "__main__":
    main()
