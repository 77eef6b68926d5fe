"""Probe: does splitting one GPU's share of the batch-8 sweep over several CUDA streams (independent model groups whose
short kernels overlap each other's ramp-up and tail) beat one grouped launch per kernel?  python tools/stream_overlap_probe.py"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from eeg_multimodal_b200 import HeadEngine  # noqa: E402

dev = torch.device("cuda:0")
DIMS, B, STEPS = (768, 768, 768), 8, 200
EPS = [0.1, 1.0, 3.0, 5.0, 8.0, 10.0]


def run(total_models, groups):
    per = total_models // groups
    engs, streams = [], []
    for gi in range(groups):
        e = HeadEngine(n_models=per, feature_dims=DIMS, eps=[EPS[(gi * per + i) % 6] for i in range(per)],
                       seeds=[980616 + (gi * per + i) // 6 for i in range(per)], precision="fp32")
        e.fast_replay = True
        engs.append(e)
        streams.append(torch.cuda.Stream())
    g = torch.Generator(device=dev).manual_seed(1)
    blocks = [torch.rand(B, d, device=dev, generator=g) for d in DIMS]
    labels = (torch.rand(B, device=dev, generator=g) < 0.66).long()
    torch.cuda.synchronize()

    def loop(n):
        for _ in range(n):
            for e, s in zip(engs, streams):
                with torch.cuda.stream(s):
                    e.train_step(blocks, labels)
    loop(10)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(STEPS)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"models": total_models, "streams": groups, "ms_per_step": round(dt / STEPS * 1e3, 4),
            "model_samples_per_s": round(total_models * B * STEPS / dt)}


if __name__ == "__main__":
    for total, groups in ((6, 1), (6, 2), (6, 3), (6, 6), (12, 1), (12, 2), (12, 4), (48, 1), (48, 2), (48, 4)):
        print(json.dumps(run(total, groups)), flush=True)
