#!/usr/bin/env python
"""Host-side cost of one HeadEngine.train_step (launch-bound regime: few models per GPU at the reference batch size).
Prints wall time per step and the cProfile top functions."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eeg_multimodal_b200 import HeadEngine  # noqa: E402


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    replay = len(sys.argv) > 3 and sys.argv[3] == "replay"
    B = 8 if precision == "fp32" else 8192
    dims = (768, 768, 768) if precision == "fp32" else (2048, 512)
    dev = torch.device("cuda:0")
    eng = HeadEngine(n_models=M, feature_dims=dims, eps=[1.0] * M, seeds=list(range(5, 5 + M)), precision=precision)
    eng.fast_replay = replay
    blocks = [torch.rand(B, d, device=dev) for d in dims]
    labels = (torch.rand(B, device=dev) < 0.66).long()
    for _ in range(5):
        eng.train_step(blocks, labels)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        eng.train_step(blocks, labels)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"M={M} {precision} replay={replay}: host enqueue {t_host / n * 1e3:.3f} ms/step, wall {t_all / n * 1e3:.3f} ms/step")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        eng.train_step(blocks, labels)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)


if __name__ == "__main__":
    main()
