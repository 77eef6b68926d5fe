"""The GPU's share of the batch-8 sweep as G independent graph-replayed step plans on G streams (M/G models each): do the
short kernels of one group fill the ramp-up and tail of the other's?  python tools/plan_groups_probe.py [M] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeg_multimodal_b200 import HeadEngine  # noqa: E402
from eeg_multimodal_b200.sweep_plan import SweepStepPlan  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
dev = torch.device("cuda:0")
dims, B, GS = (768, 768, 768), 8, 4
g = torch.Generator(device=dev).manual_seed(0)
blocks = [torch.rand(64, d, device=dev, generator=g) for d in dims]
labels = (torch.rand(64, device=dev, generator=g) < 0.66).long()
for G in (1, 2, 3, 6):
    if M % G:
        continue
    streams = [torch.cuda.Stream() for _ in range(G)]
    plans = []
    for gi in range(G):
        with torch.cuda.stream(streams[gi]):
            eng = HeadEngine(n_models=M // G, feature_dims=dims, eps=[1.0] * (M // G), seeds=list(range(gi * (M // G), (gi + 1) * (M // G))), precision="fp32")
            p = SweepStepPlan(eng, blocks, labels, B, use_pdl=True)
            p.set_rows(None)
            p.run(3)
            p.capture(GS)
            p.run(GS)
            plans.append(p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record()
    for s in streams:
        s.wait_stream(cur)
    for _ in range(steps // GS):
        for s, p in zip(streams, plans):
            with torch.cuda.stream(s):
                p.run(GS)
    for s in streams:
        cur.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"M={M} groups={G}: {ms * 1e3:.1f} us/step, {M * B / ms * 1e3:.0f} model-samples/s", flush=True)
    del plans
    torch.cuda.empty_cache()
