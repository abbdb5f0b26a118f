"""Cyclical SGHMC with full-sample storage: drop-in for methods/csghmc_fs.py (``Runner`` :17-905, ``Model`` :908-985).

Same sampler as methods/csghmc.py (fused kernel BDL_CSGHMC, Welford moments with the double-counted n), plus

  * **raw-sample store** -- in the last epochs of every cycle (``L-4 < ep % L < L-1``, ``L = epochs // num_cycles``,
    csghmc_fs.py:176) the reference dumps ``net.state_dict()`` to ``full_samples_net_ep<ep>.pth`` with a synchronous
    ``torch.save``.  Here the sample is captured into a slot of the preallocated HBM ring by ONE TMA bulk-copy launch
    (``bdl_capture_ring``, 8 B/param) and spilled to the same file asynchronously (side stream D2H + writer thread):
    the training stream never waits for PCIe or the pickler.
  * **Bayesian model average** (``evaluate_full_samples`` :260-417) -- the reference reloads every file from disk into a
    fresh ``deepcopy(net)`` and runs one pass over train / val / test per model.  Here every sample still resident in
    the ring is evaluated in place (an evaluation net whose parameters are views of the slot: no load, no copy), each
    batch crosses PCIe once and is fed to all models, and the average (``bdl_bma_mean``: fp32 running sum in sorted
    file order / model count) and the CE / error reductions (``bdl_ce_err``) run on the device.  Samples that were
    evicted from the ring, or files left in ``log_dir`` by an earlier run, are loaded from disk like the reference
    does.  Results, log lines, ``bma_evaluation_results.pkl`` and ``logits_test_bma.pkl`` keep the reference's
    keys and shapes.
  * end-of-cycle momentum reset and optional cold restarts (:129-131, :590-599).
"""
import copy
import os
import pickle
import time

import torch

from ..chain import SampleRing
from ..dist import bma_evaluate, shard_models
from ..flat import adopt_parameters, alloc_flat
from ..graphfwd import GraphedForward
from ..writer import FlatBackedStateDict
from . import csghmc as _csghmc
from ._base import reinitialize_fresh

Model = _csghmc.Model            # methods/csghmc_fs.py:908-985 is identical to methods/csghmc.py:673-781

FULL_SAMPLE_FMT = "full_samples_net_ep{ep}.pth"


class _ResidentSample:
    """One stored sample: a slot of the HBM ring + the buffers (BatchNorm statistics) the state_dict carried, and an
    evaluation network whose parameters are views of that slot."""

    def __init__(self, net, layout, row, graph=True, pool=None):
        self.row = row
        self.net = copy.deepcopy(net)
        self.net.eval()
        with torch.no_grad():
            for p, v in zip((p for _, p in self.net.named_parameters()), layout.views(row)):
                p.data = v
        for p in self.net.parameters():
            p.requires_grad_(False)
        # the slot and the buffers never move while the sample is resident: its forward is replayed as a CUDA graph
        self.forward = GraphedForward(self.net, enabled=graph, pool=pool)


class Runner(_csghmc.Runner):
    TITLE = "Cyclical SGHMC"

    def __init__(self, net, net0, args, logger):
        self.perform_cold_restarts = str(args.hparams.get("perform_cold_restarts", False)).lower() == "true"
        logger.info("Performing cold restarts: re-initializing network parameters with fresh random weights at the "
                    "start of each cycle." if self.perform_cold_restarts else
                    "Cold restarts disabled: keeping network parameters across cycles.")
        super().__init__(net, net0, args, logger)
        self.cycle_last_models_metadata = {}      # csghmc_fs.py:82-84 (kept for attribute compatibility)
        self.all_model_metadata = []
        self.model_counter = 0
        self.models_dir = os.path.join(args.log_dir, "collected_models")
        os.makedirs(self.models_dir, exist_ok=True)
        logger.info(f"Model storage directory created at: {self.models_dir}")
        self._fs_ring = None
        self._fs_resident = {}                    # file name -> _ResidentSample
        self._fs_dense = {}                       # file name -> dense view (SampleRing bookkeeping)
        self._loaders = None
        self.max_resident_samples = int(args.hparams.get("fs_ring_slots", 0)) or None

    # ---- schedule: which epochs store a raw sample (csghmc_fs.py:176) -------------------------------------
    def _stores_full_sample(self, ep):
        L = self.args.epochs // self.args.num_cycles
        return L - 4 < ep % L < L - 1

    def _expected_full_samples(self):
        return sum(1 for ep in range(self.args.epochs) if self._stores_full_sample(ep))

    # ---- training loop hooks -------------------------------------------------------------------------------
    def train(self, train_loader, val_loader, test_loader):
        self._loaders = (train_loader, val_loader, test_loader)
        return super().train(train_loader, val_loader, test_loader)

    def _after_epoch(self, ep, val_loader):
        super()._after_epoch(ep, val_loader)                 # point estimate on the validation set (:166-174)
        if self._stores_full_sample(ep):
            self.store_full_sample(ep)
            self.evaluate_full_samples(*self._loaders, desc_prefix=f"Full Samples Epoch {ep}")

    def _after_cycle_completed(self, cycle_number):          # csghmc_fs.py:590-599
        self._reset_optimizer_states()
        if self.perform_cold_restarts and cycle_number >= 1:
            self.logger.info(f"Performing COLD RESTART: Fresh random weight initialization for cycle {cycle_number + 1}")
            self._reinitialize_network_fresh()
        else:
            self.logger.info(f"Standard cycle transition: keeping weights, optimizer states reset for cycle "
                             f"{cycle_number + 1}")

    def _reset_optimizer_states(self, log=True):
        """momentum <- 0, t <- 0 (csghmc_fs.py:119-131): one memset on the flat momentum buffer."""
        if self.model.chain is not None:
            self.model.chain.reset_momenta()
        self.model.t = 0
        if log:
            self.logger.info("All optimizer states (momentum, m, v, t) reset for new cycle.")

    def _reinitialize_network_fresh(self):
        reinitialize_fresh(self.net, self.logger)

    # ---- raw-sample store -----------------------------------------------------------------------------------
    def store_full_sample(self, ep):
        """``torch.save(self.net.state_dict(), full_samples_net_ep<ep>.pth)`` (csghmc_fs.py:177): ring capture now,
        file later."""
        ch = self._chain()
        name = FULL_SAMPLE_FMT.format(ep=ep)
        path = os.path.join(self.args.log_dir, name)
        if self._fs_ring is None:
            want = self._expected_full_samples()
            if self.max_resident_samples:
                want = min(want, self.max_resident_samples)
            self._fs_ring = SampleRing(ch.layout, ch.device, max(1, want))

        def evict(old_name):                      # the slot is about to be reused: its file must be complete first
            if self._writer.is_pending(os.path.join(self.args.log_dir, old_name)):
                self._writer.flush()
            self._fs_resident.pop(old_name, None)

        self._fs_resident.pop(name, None)
        slot = self._fs_ring.capture(ch.theta, name, self._fs_dense, before_overwrite=evict)
        sample = _ResidentSample(self.net, ch.layout, self._fs_ring.buf[slot], graph=self.use_graph, pool=self._bma_graph_pool())
        self._fs_resident[name] = sample
        # on-disk contract: a state_dict (parameters from the slot, buffers from the sample's own copies)
        sd = FlatBackedStateDict.snapshot(sample.net, ch.layout, ch.names, sample.row)
        self._writer.submit(path, sd)
        return path

    def _bma_graph_pool(self):
        """One CUDA-graph memory pool for all stored models' forward graphs: they replay one after the other and every
        output is cloned at once, so they can share activation memory instead of holding S copies of it."""
        if not (self.use_graph and torch.cuda.is_available()):
            return None
        if getattr(self, "_bma_pool", None) is None:
            self._bma_pool = torch.cuda.graph_pool_handle()
        return self._bma_pool

    def _full_sample_files(self):
        """Sorted file names the reference would list (csghmc_fs.py:270): on disk or still being written."""
        log_dir = self.args.log_dir
        names = {f for f in os.listdir(log_dir) if f.startswith("full_samples_net_ep") and f.endswith(".pth")}
        names |= {n for n in self._fs_resident if self._writer.is_pending(os.path.join(log_dir, n))}
        return sorted(names)

    def _load_sample_from_disk(self, name):
        """A sample that is not resident (evicted, or written by an earlier run): load like csghmc_fs.py:300-306."""
        self._writer.flush()
        ch = self._chain()
        state = torch.load(os.path.join(self.args.log_dir, name), map_location=self.args.device)
        net = copy.deepcopy(self.net).to(self.args.device)
        flat = alloc_flat(ch.layout.n_padded, ch.device)
        adopt_parameters(net, ch.layout, flat)
        net.load_state_dict(state)
        net.eval()
        return net

    # ---- Bayesian model average ------------------------------------------------------------------------------
    def _bma_shard(self):
        """(rank, world, group) of the model-sharded BMA: hparams ``bma_shard=1`` inside an initialised process group,
        every rank calling ``evaluate_full_samples`` on the same ``log_dir``; otherwise a single rank (the reference)."""
        import torch.distributed as dist
        if int(float(self.args.hparams.get("bma_shard", 0))) and dist.is_available() and dist.is_initialized() \
                and dist.get_world_size() > 1:
            return dist.get_rank(), dist.get_world_size(), None
        return 0, 1, None

    def evaluate_full_samples(self, train_loader, val_loader, test_loader, desc_prefix="Full BMA"):
        args, logger = self.args, self.logger
        dev = args.device
        logger.info(f"Starting Bayesian Model Averaging evaluation from: {args.log_dir}")
        model_files = self._full_sample_files()
        if not model_files:
            logger.info("No model checkpoints found matching pattern 'full_samples_net_ep*.pth'.")
            return
        logger.info(f"Found {len(model_files)} model checkpoints for BMA")

        rank, world, group = self._bma_shard()
        mine = set(shard_models(len(model_files), rank, world))
        nets, used_files = [], []
        for j, name in enumerate(model_files):
            if world > 1 and j not in mine:                              # another rank's model (bma_shard=1)
                nets.append(None)
                used_files.append(name)
                continue
            res = self._fs_resident.get(name)
            if res is not None:
                nets.append(res.forward)
                used_files.append(name)
                continue
            try:
                nets.append(GraphedForward(self._load_sample_from_disk(name), enabled=self.use_graph, pool=self._bma_graph_pool()))
                used_files.append(name)
            except Exception as e:                                   # csghmc_fs.py:307-309
                if world > 1:                                        # the ranks must agree on the model list
                    raise
                logger.error(f"Failed to load model {name}: {e}")
        S = len(nets)
        my_nets = {j: net for j, net in enumerate(nets) if net is not None}

        bma_results = {}
        for dataset_name, loader in (("train", train_loader), ("val", val_loader), ("test", test_loader)):
            if loader is None:
                continue
            logger.info(f"Performing BMA evaluation on {dataset_name} set...")
            if S == 0:
                logger.warning(f"No valid models found for {dataset_name} evaluation")
                continue
            tic = time.time()
            r = bma_evaluate(my_nets, S, loader, dev, rank=rank, world=world, group=group)
            nb, loss_per, err_per = r["n"], r["loss_per"], r["err_per"]
            for name, ls, er in zip(used_files, loss_per, err_per):
                logger.info(f"Model {name} on {dataset_name}: loss={ls / nb:.4f}, error={er / nb:.4f}")
            bma_loss, bma_error = r["bma_loss_sum"] / nb, r["bma_err_sum"] / nb
            bma_results[dataset_name] = {
                "loss": bma_loss, "error": bma_error, "num_models": S,
                "targets": r["targets"], "logits": r["logits"],
                "logits_all": r["logits_all"],                            # [samples, classes, models]
                # reference quirk kept: total_samples is accumulated once per model, so these two are divided by
                # nb * S * S, not nb * S (csghmc_fs.py:356-358, 385-386)
                "individual_avg_loss": loss_per.sum() / (nb * S * S),
                "individual_avg_error": err_per.sum() / (nb * S * S),
            }
            logger.info(f"BMA results on {dataset_name}: loss={bma_loss:.4f}, error={bma_error:.4f} "
                        f"(averaged over {S} models; {time.time() - tic:.3f} s)")
            logger.info(f"Individual models average on {dataset_name}: "
                        f"loss={bma_results[dataset_name]['individual_avg_loss']:.4f}, "
                        f"error={bma_results[dataset_name]['individual_avg_error']:.4f}")

        if rank != 0:                                                     # sharded BMA: rank 0 owns the files
            return bma_results
        results_path = os.path.join(args.log_dir, "bma_evaluation_results.pkl")
        with open(results_path, "wb") as f:
            pickle.dump(bma_results, f)
        logger.info(f"BMA evaluation results saved to {results_path}")
        if "test" in bma_results:
            t = bma_results["test"]
            fname = self.save_logits(t["targets"], t["logits"], t["logits_all"], suffix="test_bma")
            logger.info(f"BMA test predictions saved at {fname}")
        logger.info("Finished Bayesian Model Averaging evaluation.")
        return bma_results
