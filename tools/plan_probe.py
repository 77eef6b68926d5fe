"""A few fused sweep steps (direct launches or graph) for ncu / timing probes.  python tools/plan_probe.py [M] [steps] [graph_steps] [pdl]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeg_multimodal_b200 import HeadEngine  # noqa: E402
from eeg_multimodal_b200.sweep_plan import SweepStepPlan  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
gsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
pdl = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
dev = torch.device("cuda:0")
dims, B = (768, 768, 768), 8
eng = HeadEngine(n_models=M, feature_dims=dims, eps=[1.0] * M, seeds=list(range(M)), precision="fp32")
g = torch.Generator(device=dev).manual_seed(0)
blocks = [torch.rand(64, d, device=dev, generator=g) for d in dims]
labels = (torch.rand(64, device=dev, generator=g) < 0.66).long()
plan = SweepStepPlan(eng, blocks, labels, B, use_pdl=pdl)
plan.set_rows(None)
plan.run(3)
if gsteps:
    plan.capture(gsteps)
    plan.run(gsteps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
plan.run(steps)
e1.record()
torch.cuda.synchronize()
print(f"M={M} steps={steps} graph_steps={gsteps} pdl={pdl}: {e0.elapsed_time(e1) / steps * 1e3:.1f} us/step")
