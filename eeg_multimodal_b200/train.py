"""`train.py`-compatible driver for the privatised fusion head.

Keeps the reference's flags (train.py:29-47: --exp --name --batch_size --data_name --eps --n_class
--n_dp --n_para --n_eval --n_epochs --interval --metrics) and its outputs (`experiment/<exp>/<name>/
{debug.log, info.log, model.pth, results.pth}`, model.pth saved on best mean accuracy,
train.py:139-144), with the encoders replaced by a feature cache (feature_cache.py) and the model
replaced by the HeadEngine, which also trains a whole eps x seed sweep at once and shards it over
ranks (`torchrun --nproc-per-node N -m eeg_multimodal_b200.train --eps-list 0.1,1,3,5,8,10 --n-seeds 8`).

Differences that are deliberate: `--n_dp` defaults to 1 = the two-pass step of past_acc.py:198-212
(train.py has the DP pass commented out, train.py:100-105; `--n_dp 0` reproduces that);
`--metrics` supports Accuracy and F1Score without torchmetrics (absent in this image).
`--n_eval` repeats every evaluation batch with fresh noise (train.py:126-131).
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
from .feature_cache import FeatureLoader, load_features, synthetic_features
from .records import EpochMeter, RecordWriter, binary_f1


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
    p.add_argument("--precision", choices=["fp32", "bf16"], default="fp32")
    p.add_argument("--unfixed-formula", action="store_true", help="eps_hat = log(..) as in model.py:57 (new_*eps runs)")
    p.add_argument("--records-root", type=str, default=None, help="also write model_dict-style records per model")
    return p


def run(cfg) -> dict:
    if cfg.n_class != 2:
        raise NotImplementedError("the reference head is binary (nn.Linear(768, 2))")
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
    train_loader = FeatureLoader(tb, tl, cfg.batch_size, shuffle=True, seed=980616, device=dev)
    val_loader = FeatureLoader(vb, vl, cfg.batch_size, shuffle=True, seed=980616, device=dev)

    eps_list = [float(x) for x in cfg.eps_list.split(",")] if cfg.eps_list else [cfg.eps]
    variants = [v.strip() or None for v in cfg.variants.split(",")] if cfg.variants else [None]
    grid = parallel.sweep_grid(eps_list, cfg.n_seeds, variants=tuple(variants))
    mine = [grid[i] for i in parallel.shard_models(len(grid), world, rank)]
    if not mine:
        return {}
    fmean = None
    if any(v and (v.startswith("newinit_") or v == "feawei") for v in variants):
        fmean = dp_variants.feature_mean(tb)                  # what the reference keeps in feawei.pkl
    dp0 = [dp_variants.dp_init(g["variant"], dims, fmean) for g in mine]
    eng = HeadEngine(n_models=len(mine), feature_dims=dims, eps=[g["eps"] for g in mine], seeds=[g["seed"] for g in mine],
                     lr=cfg.lr, precision=cfg.precision, fixed_formula=not cfg.unfixed_formula,
                     init_seed=980616, device=dev, dp_init=dp0)
    writers = [RecordWriter(cfg.records_root, f"newfrac_{g['eps']}eps" + (f"_{g['variant']}" if g["variant"] else "") + f"_seed{g['seed']}/")
               for g in mine] if cfg.records_root else None
    want = [m.strip() for m in cfg.metrics.split(",")]
    results = {"Accuracy": [], "F1Score": [], "val_loss": [], "train_loss": []}
    best_acc = [0.0] * len(mine)
    for epoch in range(cfg.n_epochs):
        tm = [EpochMeter() for _ in mine]
        for i, (blocks, labels) in enumerate(train_loader):
            for _ in range(cfg.n_para):
                st = eng.train_step(blocks, labels, dp_pass=cfg.n_dp > 0)
            loss, acc = st["loss"].tolist(), st["acc"].tolist()
            for k in range(len(mine)):
                tm[k].update(loss[k], acc[k])
            logger.debug(f"Train Epoch: {epoch:3d} [{i + 1:3d}/{len(train_loader):3d}] loss {sum(loss) / len(loss):.4f}")
        results["train_loss"].append([m.loss for m in tm])
        if (epoch + 1) % cfg.interval:
            continue
        vm = [EpochMeter() for _ in mine]
        accs = torch.zeros(len(mine), cfg.n_eval)
        f1s = torch.zeros(len(mine), cfg.n_eval)
        # everything stays on the device while the evaluation launches are queued; ONE transfer per epoch
        preds = [[] for _ in range(cfg.n_eval)]                        # per repetition: list of [M,B] predictions
        labs, first_stats = [], []
        for blocks, labels in val_loader:
            labs.append(labels.reshape(-1))
            for e in range(cfg.n_eval):                               # train.py:126-131
                ev = eng.eval_step(blocks, labels)
                preds[e].append(ev["pred"].view(len(mine), -1).clone())
                if e == 0:
                    first_stats.append(torch.stack((ev["loss"], ev["acc"]), dim=1).clone())
        lab = torch.cat(labs).cpu()
        preds = [torch.cat(p, dim=1).cpu() for p in preds]           # n_eval x [M, N]
        first_stats = torch.stack(first_stats).cpu().tolist()         # [batches][M][2]
        lo = 0
        for bi, l in enumerate(labs):                                 # epoch meters: unweighted per-batch means (past_acc.py:236)
            hi = lo + l.numel()
            for k in range(len(mine)):
                vm[k].update(first_stats[bi][k][0], first_stats[bi][k][1], preds[0][k, lo:hi], lab[lo:hi])
            lo = hi
        for k in range(len(mine)):
            for e in range(cfg.n_eval):
                accs[k, e] = (preds[e][k] == lab).float().mean()
                f1s[k, e] = binary_f1(preds[e][k], lab)
        info = f"Eval  Epoch: {epoch:3d}"
        if "Accuracy" in want:
            info += f" | Accuracy: {accs.mean().item():5.2f}"
        if "F1Score" in want:
            info += f" | F1Score: {f1s.mean().item():5.2f}"
        logger.info(info)
        results["Accuracy"].append(accs.clone())
        results["F1Score"].append(f1s.clone())
        results["val_loss"].append([m.loss for m in vm])
        for k in range(len(mine)):
            if accs[k].mean() > best_acc[k]:                          # train.py:139-143
                best_acc[k] = float(accs[k].mean())
                if rank == 0 and k == 0:
                    torch.save(eng.state_dict(0), os.path.join(base, "model.pth"))
            if writers:
                writers[k].epoch_end(epoch + 1, tm[k], vm[k], lambda k=k: eng.state_dict(k))
    out = {"grid": mine, "best_acc": best_acc, "Accuracy": results["Accuracy"], "F1Score": results["F1Score"],
           "train_loss": results["train_loss"], "val_loss": results["val_loss"], "DP_params": eng.DP.detach().cpu()}
    merged = parallel.gather_metrics({g["index"]: best_acc[k] for k, g in enumerate(mine)})
    if rank == 0:
        torch.save(out, os.path.join(base, "results.pth"))
        logger.info(f"best accuracy per model: {merged}")
    out["best_acc_all"] = merged
    return out


def main(argv=None):
    import torch.distributed as dist

    try:
        run(build_parser().parse_args(argv))
    finally:
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
