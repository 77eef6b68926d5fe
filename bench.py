#!/usr/bin/env python
"""bench.py -- model-samples/s of the PriGumbel eps x seed sweep on B200 (BASELINE.json metric).

Default workload (`sweep_synth64k`): every GPU trains `--models-per-gpu` independent heads of the
sweep (eps in {0.1,1,3,5,8,10} x seeds; 6 per GPU == the 48-model sweep on 8 GPUs) on the
synthetic shape of BASELINE config 4: batch 65,536 x (2048 EEG + 512 action) fp32 features,
D=2560 -> 2560 -> 768 -> 2, bf16 tensor-core GEMMs with fp32 accumulate and fp32 Adam.  One
"step" = one full reference step (past_acc.py:198-212: hard=False pass + Adam(DP), hard=True pass
+ Adam(weights)) of every model of the rank on one batch.  Models are independent, so ranks share
nothing: weak scaling, no collective on the data path (`--workload dp64k` is the single-model
data-parallel mode with an NCCL all-reduce of the gradients).

  value  samples x models / s with the batches resident in HBM (8 resident batches of 671 MB,
         cycled: every step reads inputs far larger than the 126 MB L2)
  e2e    same metric through the public API with HOST (pinned) feature buffers: per step one H2D
         copy of the batch (double-buffered on a copy stream) and one D2H read of the losses
  --impl reference   the reference's CPU PyTorch path (oracle restatement with the reference's own
         noise calls), on all host cores, on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EPS_SET = [0.1, 1.0, 3.0, 5.0, 8.0, 10.0]
DIMS = (2048, 512)
HIDDEN = 768
BATCH = 65536


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sweep_synth64k", choices=["sweep_synth64k", "dp64k", "sweep48_b8"])
    ap.add_argument("--models-per-gpu", type=int, default=6)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--resident-batches", type=int, default=8)
    ap.add_argument("--cpu-sample", type=int, default=2048, help="samples per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def flops_per_sample_step(D, H):
    """Tensor-pipe flops of one reference step per sample with the discarded gradients skipped:
    pass 1: fwd (D^2 + DH) + dH1 (DH) + dX (D^2);  pass 2: fwd + dW2 + dH1 + dW1."""
    return 2 * ((D * D + D * H) + D * H + D * D) + 2 * ((D * D + D * H) + D * H + D * H + D * D)


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# --------------------------------------------------------------------------------------------------
def cpu_reference_leg(args, steps, warmup, sample):
    """The reference's CPU path (oracle restatement incl. the reference's own host noise calls),
    all host threads, on a bounded sample of the workload.  Returns (samples/s, cores, sample str)."""
    import torch

    from oracle import head_oracle as ho

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D = sum(DIMS)
    torch.manual_seed(980616)
    blocks = [torch.rand(sample, d) for d in DIMS]
    label = (torch.rand(sample, 1) < 0.66).long()
    p = ho.make_params(D, HIDDEN, seed=0).clone(requires_grad=True)
    for _ in range(max(1, warmup)):
        ho.reference_step_cpu(blocks, label, p, 1.0)
    t0 = time.perf_counter()
    for _ in range(steps):
        ho.reference_step_cpu(blocks, label, p, 1.0)
    dt = time.perf_counter() - t0
    return sample * steps / dt, cores, f"{steps} two-pass fwd+bwd steps of one model on {sample} samples x D={D} (fp32, torch CPU)", dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, cores, sample, sec = cpu_reference_leg(args, args.steps, args.warmup, args.cpu_sample)
    line = {"impl": "reference", "metric": "model-samples/sec", "value": v, "unit": "model-samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "note": "reference CPU path; bounded sample of the same workload"},
            "cpu_baseline": {"value": v, "unit": "model-samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "model-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from eeg_multimodal_b200 import HeadEngine, _lib, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    K, W = args.steps, max(3, args.warmup)
    B, D = args.batch, sum(DIMS)

    if args.workload == "sweep48_b8":
        dims, B, M, precision = (768, 768, 768), 8, args.models_per_gpu, "fp32"
        D = 2304
    elif args.workload == "dp64k":
        dims, M, precision = DIMS, 1, "bf16"
        B = args.batch // world
    else:
        dims, M, precision = DIMS, args.models_per_gpu, "bf16"
    eps = [EPS_SET[(rank * M + i) % len(EPS_SET)] for i in range(M)]
    seeds = [980616 + (rank * M + i) // len(EPS_SET) for i in range(M)] if args.workload != "dp64k" else [980616]
    eng = HeadEngine(n_models=M, feature_dims=dims, hidden=HIDDEN, eps=eps, seeds=seeds, precision=precision,
                     init_seed=980616 + 1000 * (rank if args.workload != "dp64k" else 0))

    # ---- synthetic data resident in HBM (U(0,1) features, Bernoulli(0.66) labels)
    g = torch.Generator(device=dev).manual_seed(980616 + rank)
    nres = max(1, args.resident_batches)
    data = [([torch.rand(B, d, device=dev, generator=g) for d in dims], (torch.rand(B, device=dev, generator=g) < 0.66).long())
            for _ in range(nres)]
    global_batch = B * world if args.workload == "dp64k" else None
    hook = None
    if args.workload == "dp64k" and world > 1:
        def hook(t):
            dist.all_reduce(t)

    def step(i):
        blocks, labels = data[i % nres]
        row0 = (i * B * world + rank * B) if args.workload == "dp64k" else i * B
        return eng.train_step(blocks, labels, row0=row0, global_batch=global_batch, grad_hook=hook)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(W):
        step(i)
    fence()
    sampler = ClockSampler(local)
    sampler.start()
    _lib.launch_count = 0
    ops.GEMM_TIMING = [] if precision == "bf16" else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        st = step(W + i)
    e1.record()
    fence()
    launches = _lib.launch_count
    gemm_t = ops.GEMM_TIMING
    ops.GEMM_TIMING = None
    ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    tms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms)
    total_models = M * world if args.workload != "dp64k" else 1
    samples_per_step = total_models * B * (world if args.workload == "dp64k" else 1)
    value = samples_per_step * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel (largest share of GEMM time), from the timed region's own events
    peaks, peak_kind = measured_peaks()
    roofline = None
    if gemm_t:
        agg = {}
        for tag, a, b in gemm_t:
            t, n = agg.get(tag, (0.0, 0))
            agg[tag] = (t + a.elapsed_time(b), n + 1)
        tag, (t_ms, n) = max(agg.items(), key=lambda kv: kv[1][0])
        m_, n_, k_ = tag[:3]
        achieved = 2.0 * m_ * n_ * k_ / (t_ms / n * 1e-3) / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        roofline = {"bound": "tensor", "kernel": f"gemm_bf16_tc M={m_} N={n_} K={k_} a_mn={tag[3]} b_mn={tag[4]} epi={tag[5]}",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "peak_kind": f"{peak_kind} (sustained cuBLAS bf16)", "traffic": None,
                    "share_of_step": t_ms / ms, "avg_launch_ms": t_ms / n,
                    "all_gemms_ms_per_step": sum(v[0] for v in agg.values()) / K,
                    "step_tensor_frac": flops_per_sample_step(D, HIDDEN) * B * M * K / (ms * 1e-3) / 1e12 / peak}

    # ---- end to end through the public API: host (pinned) buffers, H2D per step, D2H of the losses
    e2e = None
    if not args.no_e2e:
        host = [([torch.rand(B, d).pin_memory() for d in dims], (torch.rand(B) < 0.66).long().pin_memory()) for _ in range(2)]
        devbuf = [([torch.empty(B, d, device=dev) for d in dims], torch.empty(B, dtype=torch.int64, device=dev)) for _ in range(2)]
        copy_stream = torch.cuda.Stream()
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        loss_host = torch.empty(M, 4).pin_memory()

        def upload(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                for hb, db in zip(host[s][0], devbuf[s][0]):
                    db.copy_(hb, non_blocking=True)
                devbuf[s][1].copy_(host[s][1], non_blocking=True)
                ready[s].record(copy_stream)

        def e2e_loop(n, first):
            for s in range(2):
                freed[s].record()
            upload(first)
            for i in range(first, first + n):
                s = i % 2
                if i + 1 < first + n:
                    upload(i + 1)
                torch.cuda.current_stream().wait_event(ready[s])
                st = eng.train_step(devbuf[s][0], devbuf[s][1], row0=i * B, global_batch=global_batch, grad_hook=hook)
                freed[s].record()
                loss_host[:, :3].copy_(torch.stack([st["loss"], st["n_correct"], st["acc"]], 1), non_blocking=True)
            torch.cuda.synchronize()

        e2e_loop(2, 0)
        fence()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_loop(K, 0)
        t1.record()
        fence()
        ems = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        h2d = sum(B * d * 4 for d in dims) + B * 8
        e2e = {"value": samples_per_step * K / (float(ems) * 1e-3), "unit": "model-samples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": M * 3 * 4, "ms_per_step": float(ems) / K}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference_leg(args, 3, 1, args.cpu_sample)
        cpu = {"value": v, "unit": "model-samples/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {"metric": "model-samples/sec", "value": value, "unit": "model-samples/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if args.workload == "dp64k" else "weak",
                "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": args.workload, "models_per_gpu": M, "models_total": total_models, "batch_per_model": B,
                           "feature_dims": list(dims), "hidden": HIDDEN, "eps": eps, "step": "reference two-pass step incl. both Adam updates",
                           "l2": f"{nres} resident batches of {sum(dims) * B * 4 / 1e6:.0f} MB cycled (inputs >> 126 MB L2)",
                           "parallelism": "independent models per GPU, no collective" if args.workload != "dp64k" else f"dp{world} NCCL all-reduce"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(),
                "loss_last": [float(x) for x in st["loss"].cpu()]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
