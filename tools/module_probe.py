"""Forward + backward of the module API (ConcatModel.forward -> cal_loss -> backward) at a large batch, per C-ABI call.
python tools/module_probe.py [precision] [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeg_multimodal_b200 import ConcatModel, cal_loss, ops  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
dev = torch.device("cuda:0")
dims = (2048, 512)
model = ConcatModel(feature_dims=dims, precision=prec).to(dev)
model.eps = torch.tensor(1.0)
g = torch.Generator(device=dev).manual_seed(0)
x = tuple(torch.rand(B, d, device=dev, generator=g) for d in dims)
y = (torch.rand(B, 1, device=dev, generator=g) < 0.66).long()


def step():
    model.zero_grad(set_to_none=True)
    loss, acc, _, _ = cal_loss(model(x, hard=True), y)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"ConcatModel precision={prec} B={B}: {ms:.2f} ms per forward+backward = {B / ms * 1e3 / 1e6:.2f} M samples/s, loss {float(loss):.4f}")
ops.TIMING = []
step()
torch.cuda.synchronize()
agg = {}
for tag, a, b in ops.TIMING:
    k = " ".join(str(t) for t in tag[:5])
    agg[k] = agg.get(k, 0.0) + a.elapsed_time(b)
ops.TIMING = None
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]:
    print(f"  {v:8.3f} ms  {k}")
