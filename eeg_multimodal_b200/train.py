"""`train.py`-compatible driver for the privatised fusion head.

Keeps the reference's flags (train.py:29-47: --exp --name --batch_size --data_name --eps --n_class
--n_dp --n_para --n_eval --n_epochs --interval --metrics) and its outputs (`experiment/<exp>/<name>/
{debug.log, info.log, model.pth, results.pth}`, model.pth saved on best mean accuracy,
train.py:139-144), with the encoders replaced by a feature cache (feature_cache.py) and the model
replaced by the HeadEngine, which also trains a whole eps x seed sweep at once and shards it over
ranks (`torchrun --nproc-per-node N -m eeg_multimodal_b200.train --eps-list 0.1,1,3,5,8,10 --n-seeds 8`).

Nothing in the training loop runs on the host: both splits are resident in HBM (ResidentDataset), an
epoch's shuffle is one device permutation, at the reference's batch sizes (<= 8, fp32) every step is
one CUDA-graph replay of the fused sweep step (sweep_plan.SweepStepPlan), the `n_eval` repeated
stochastic evaluations of a batch (train.py:126-131) are one batched pass with n_eval Philox offsets,
and statistics cross to the host once per epoch.

results.pth: the reference's keys and shapes for grid model 0 (logits [E*N, n_eval, 2], pred, val_loss,
train_loss, Accuracy, DP_params -- see reference_results) plus `sweep` / `sweep_results` with the same
for every model of the grid; `model_<index>.pth` per model (model.pth = model 0), gathered over ranks.

Differences that are deliberate: `--n_dp` defaults to 1 = the two-pass step of past_acc.py:198-212
(train.py has the DP pass commented out, train.py:100-105; `--n_dp 0` reproduces that);
`--n_para` / `--n_dp` > 1 (gradient accumulation over repeated noisy passes) are refused, not
reinterpreted; `--metrics` supports Accuracy and F1Score without torchmetrics (absent in this image).
"""
from __future__ import annotations

import argparse
import logging
import os
import sys

import torch

from . import parallel
from . import variants as dp_variants
from .engine import HeadEngine
from .feature_cache import ResidentDataset, load_features, synthetic_features
from .records import EpochMeter, RecordWriter, binary_f1
from .sweep_plan import SweepStepPlan


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--exp", type=str, default="test")
    p.add_argument("--name", type=str, default="test")
    p.add_argument("--batch_size", "-bs", type=int, default=8)
    p.add_argument("--data_name", "-d", type=str, default="EEG")
    p.add_argument("--eps", "-e", type=float, default=2.0)
    p.add_argument("--n_class", "-c", type=int, default=2)
    p.add_argument("--n_dp", "-nd", type=int, default=1)
    p.add_argument("--n_para", "-np", type=int, default=1)
    p.add_argument("--n_eval", "-ne", type=int, default=5)
    p.add_argument("--n_epochs", "-n", type=int, default=50)
    p.add_argument("--interval", type=int, default=1)
    p.add_argument("--metrics", "-m", type=str, default="Accuracy")
    # additions
    p.add_argument("--features", type=str, default=None, help="train feature cache (.npz)")
    p.add_argument("--val-features", type=str, default=None)
    p.add_argument("--synthetic", type=int, default=0, help="use N synthetic samples (U(0,1), Bernoulli(0.66))")
    p.add_argument("--feature-dims", type=str, default="768,768,768")
    p.add_argument("--eps-list", type=str, default=None, help="comma list: train a sweep instead of one model")
    p.add_argument("--n-seeds", type=int, default=1)
    p.add_argument("--variants", type=str, default=None,
                   help="comma list of DP initialisation variants (zeros,newinit,tt,newinit_k1,newinit_k3,feawei): model_dict/newfrac_*")
    p.add_argument("--lr", type=float, default=1e-6)        # past_acc.py:157, train.py:75
    p.add_argument("--precision", choices=["fp32", "bf16", "fp32x3"], default="fp32",
                   help="arithmetic of the dense layers: fp32 = the reference's (CUDA-core kernels at the reference batch sizes, the "
                        "fp32x3 tensor-core route from 1,024 rows on), fp32x3 = always that route, bf16 = bf16 GEMM operands")
    p.add_argument("--unfixed-formula", action="store_true", help="eps_hat = log(..) as in model.py:57 (new_*eps runs)")
    p.add_argument("--records-root", type=str, default=None, help="also write model_dict-style records per model")
    return p


def _reject_unsupported(cfg):
    """Flags whose reference semantics this driver does not implement are refused rather than silently reinterpreted.
    train.py:108-112 accumulates the gradients of `n_para` noisy passes of loss.sum() into ONE Adam step (and, in the
    commented-out DP pass, of `n_dp` passes): the engine applies one update per pass on the mean loss."""
    if cfg.n_para != 1:
        raise NotImplementedError("--n_para > 1 (gradient accumulation over repeated noisy passes, train.py:108-112) is not implemented")
    if cfg.n_dp not in (0, 1):
        raise NotImplementedError("--n_dp > 1 (accumulated DP passes, train.py:100-105) is not implemented; 0 = no DP pass, 1 = one")
    if cfg.n_class != 2:
        raise NotImplementedError("the reference head is binary (nn.Linear(768, 2))")


def reference_results(epochs, n_eval):
    """Per-epoch records of ONE model -> the dict the reference saves as results.pth (train.py:131-144): every value a
    torch.cat over epochs.  logits [E*N, n_eval, 2], pred [E*N, n_eval], val_loss [E*N, n_eval] (per-sample CE),
    train_loss [steps*B] (per-sample CE of every training pass), <metric> [E*n_eval], DP_params [E, D].
    Layout note: the reference builds the [N, n_eval, ...] tensors with `torch.cat(v).view(-1, n_eval, ...)` over a list
    ordered (batch, repetition, row), which interleaves rows of one repetition; here axis 1 really is the repetition."""
    out = {}
    for k in ("logits", "pred", "val_loss", "train_loss", "Accuracy", "F1Score", "DP_params"):
        vals = [e[k] for e in epochs if k in e]
        if vals:
            out[k] = torch.cat(vals)
    return out


def run(cfg) -> dict:
    _reject_unsupported(cfg)
    rank, world = parallel.init_distributed()
    base = f"experiment/{cfg.exp}/{cfg.name}/"
    os.makedirs(base, exist_ok=True)
    logger = logging.getLogger(f"pgfuse.train.{rank}")
    logger.setLevel(logging.DEBUG)
    logger.handlers.clear()
    if rank == 0:
        fmt = logging.Formatter("%(asctime)s - %(levelname)s - %(message)s")
        for h, lvl in ((logging.FileHandler(base + "debug.log", "w"), logging.DEBUG),
                       (logging.FileHandler(base + "info.log", "w"), logging.INFO), (logging.StreamHandler(sys.stdout), logging.INFO)):
            h.setLevel(lvl)
            h.setFormatter(fmt)
            logger.addHandler(h)
    logger.info(cfg)

    eps_list = [float(x) for x in cfg.eps_list.split(",")] if cfg.eps_list else [cfg.eps]
    variants = [v.strip() or None for v in cfg.variants.split(",")] if cfg.variants else [None]
    grid = parallel.sweep_grid(eps_list, cfg.n_seeds, variants=tuple(variants))
    mine = [grid[i] for i in parallel.shard_models(len(grid), world, rank)]
    local = _train_models(cfg, mine, variants, rank, base, logger) if mine else {}
    # every rank -- also one that owns no model (world > grid size) -- takes part in the gathers
    merged = parallel.gather_metrics({i: r["best_acc"] for i, r in local.items()})
    everything = parallel.gather_results(local)
    if rank == 0:
        first = min(everything) if everything else None
        ref = dict(everything[first]["reference"]) if first is not None else {}
        ref["sweep"] = {i: {k: v for k, v in r.items() if k != "reference"} for i, r in everything.items()}
        ref["sweep_results"] = {i: r["reference"] for i, r in everything.items()}
        torch.save(ref, os.path.join(base, "results.pth"))
        logger.info(f"best accuracy per model: {merged}")
    return {"grid": mine, "local": local, "best_acc_all": merged, "results": everything if rank == 0 else None}


def _train_models(cfg, mine, variants, rank, base, logger) -> dict:
    """Train this rank's share of the grid.  Returns {grid index: dict(best_acc, eps, seed, variant, reference=...)}."""
    dims = tuple(int(x) for x in cfg.feature_dims.split(","))
    if cfg.features:
        tb, tl = load_features(cfg.features)
        vb, vl = load_features(cfg.val_features) if cfg.val_features else (tb, tl)
        dims = tuple(b.shape[1] for b in tb)
    else:
        n = cfg.synthetic or 64 * cfg.batch_size
        tb, tl = synthetic_features(n, dims, seed=980616)
        vb, vl = synthetic_features(max(cfg.batch_size, n // 4), dims, seed=980617)
    dev = torch.device("cuda", torch.cuda.current_device())
    # both splits live in HBM; an epoch's shuffle is one device permutation (data.py:41-42: shuffle=True for train AND val)
    train = ResidentDataset(tb, tl, cfg.batch_size, shuffle=True, seed=980616, device=dev)
    val = ResidentDataset(vb, vl, cfg.batch_size, shuffle=True, seed=980616, device=dev)
    fmean = None
    if any(v and (v.startswith("newinit_") or v == "feawei") for v in variants):
        fmean = dp_variants.feature_mean(tb)                  # what the reference keeps in feawei.pkl
    dp0 = [dp_variants.dp_init(g["variant"], dims, fmean) for g in mine]
    M = len(mine)
    eng = HeadEngine(n_models=M, feature_dims=dims, eps=[g["eps"] for g in mine], seeds=[g["seed"] for g in mine],
                     lr=cfg.lr, precision=cfg.precision, fixed_formula=not cfg.unfixed_formula,
                     init_seed=980616, device=dev, dp_init=dp0)
    writers = [RecordWriter(cfg.records_root, f"newfrac_{g['eps']}eps" + (f"_{g['variant']}" if g["variant"] else "") + f"_seed{g['seed']}/")
               for g in mine] if cfg.records_root else None
    want = [m.strip() for m in cfg.metrics.split(",")]
    bs, n_eval = cfg.batch_size, cfg.n_eval
    # the reference's batch sizes on the fp32 path: the whole step is one graph replay per batch, gathered on the device
    plan = None
    if cfg.precision == "fp32" and bs <= 8 and train.n_full > 0:
        plan = SweepStepPlan(eng, train.blocks, train.labels, bs, dp_pass=cfg.n_dp > 0, wrap=False)
    epochs = [[] for _ in mine]
    best_acc = [0.0] * M
    for epoch in range(cfg.n_epochs):
        order = train.epoch_order()
        stats, logit_log = [], []          # device tensors; ONE transfer per epoch
        n_full = train.n_full
        if plan is not None:
            plan.set_rows(order, cursor=0)
            if not plan._captured:
                plan.run(1)
                stats.append(plan.stats_model.clone())
                logit_log.append((plan.logits.clone(), train.labels.index_select(0, order[:bs])))
                plan.capture(1)
                first = 1
            else:
                first = 0
            for i in range(first, n_full):
                plan.run(1)
                stats.append(plan.stats_model.clone())
                logit_log.append((plan.logits.clone(), train.labels.index_select(0, order[i * bs:(i + 1) * bs])))
            tail = range(n_full, len(train))
        else:
            tail = range(len(train))
        for i in tail:                     # the partial last batch (DataLoader keeps it), or every batch off the plan path
            blocks, labels = train.batch(i)
            st = eng.train_step(blocks, labels, dp_pass=cfg.n_dp > 0)
            stats.append(st["stats"])
            logit_log.append((eng._buf("logits_model", (M, labels.shape[0], 2), torch.float32).clone(), labels))
        stats = torch.stack(stats).cpu()                              # [batches, M, 4]
        tm = [EpochMeter() for _ in mine]
        for bi in range(stats.shape[0]):
            for k in range(M):
                tm[k].update(float(stats[bi, k, 0]), float(stats[bi, k, 2]))
        # per-sample training losses, as train.py:110-111 records them (criterion(reduction='none'))
        tr_loss = [torch.cat([torch.nn.functional.cross_entropy(lg[k], lab, reduction="none") for lg, lab in logit_log]).cpu()
                   for k in range(M)]
        logger.debug(f"Train Epoch: {epoch:3d} loss {float(stats[:, :, 0].mean()):.4f}")
        rec = [{"train_loss": tr_loss[k]} for k in range(M)]
        if (epoch + 1) % cfg.interval == 0:
            val.epoch_order()
            preds, logits, labs, first_stats = [], [], [], []
            for i in range(len(val)):                                 # n_eval repetitions = ONE batched pass (train.py:126-131)
                blocks, labels = val.batch(i)
                ev = eng.eval_step(blocks, labels, n_eval=n_eval)
                preds.append(ev["pred"].reshape(M, n_eval, -1))
                logits.append(ev["logits"].reshape(M, n_eval, -1, 2))
                labs.append(labels)
            lab_d = torch.cat(labs)
            pred_d = torch.cat(preds, dim=2)                          # [M, n_eval, N]
            logit_d = torch.cat(logits, dim=2)                        # [M, n_eval, N, 2]
            vloss_d = torch.nn.functional.cross_entropy(logit_d.reshape(-1, 2), lab_d.repeat(M * n_eval), reduction="none").view(M, n_eval, -1)
            lab, pred, logit, vloss = lab_d.cpu(), pred_d.cpu(), logit_d.cpu(), vloss_d.cpu()
            accs = (pred == lab).float().mean(dim=2)                  # [M, n_eval]
            f1s = torch.tensor([[binary_f1(pred[k, e], lab) for e in range(n_eval)] for k in range(M)])
            vm = [EpochMeter() for _ in mine]
            lo = 0
            for l in labs:                                            # epoch meters: unweighted per-batch means (past_acc.py:236)
                hi = lo + l.numel()
                for k in range(M):
                    vm[k].update(float(vloss[k, 0, lo:hi].mean()), float((pred[k, 0, lo:hi] == lab[lo:hi]).float().mean()),
                                 pred[k, 0, lo:hi], lab[lo:hi])
                lo = hi
            info = f"Eval  Epoch: {epoch:3d}"
            if "Accuracy" in want:
                info += f" | Accuracy: {accs.mean().item():5.2f}"
            if "F1Score" in want:
                info += f" | F1Score: {f1s.mean().item():5.2f}"
            logger.info(info)
            dp_now = eng.DP.detach().cpu()
            for k, g in enumerate(mine):
                rec[k].update(logits=logit[k].permute(1, 0, 2).contiguous(), pred=pred[k].t().contiguous(),
                              val_loss=vloss[k].t().contiguous(), Accuracy=accs[k].clone(), F1Score=f1s[k].clone(),
                              DP_params=dp_now[k].view(1, -1))
                if accs[k].mean() > best_acc[k]:                      # train.py:139-143, one checkpoint per model
                    best_acc[k] = float(accs[k].mean())
                    torch.save(eng.state_dict(k), os.path.join(base, f"model_{g['index']}.pth"))
                    if g["index"] == 0:
                        torch.save(eng.state_dict(k), os.path.join(base, "model.pth"))
                if writers:
                    writers[k].epoch_end(epoch + 1, tm[k], vm[k], lambda k=k: eng.state_dict(k))
        for k in range(M):
            epochs[k].append(rec[k])
    if plan is not None:
        plan.close()
    return {g["index"]: dict(best_acc=best_acc[k], eps=g["eps"], seed=g["seed"], variant=g["variant"],
                             reference=reference_results(epochs[k], n_eval)) for k, g in enumerate(mine)}


def main(argv=None):
    import torch.distributed as dist

    try:
        run(build_parser().parse_args(argv))
    finally:
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
