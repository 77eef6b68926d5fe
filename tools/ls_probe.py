"""Time the TMA-ring forward kernel alone (M models, W1 shape).  Env knobs: PGF_LS_DBG, PGF_LS_STAGES."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeg_multimodal_b200 import ops, _lib as L  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 6
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2304
K = 2304
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
# several weight sets so that consecutive launches do not hit L2
sets = [torch.randn(M, N, K, device=dev, generator=g) * 0.02 for _ in range(max(2, 600 // (M * N * K * 4 // 2 ** 20 + 1)))][:8]
X = torch.rand(M, 8, K, device=dev, generator=g)
b = torch.zeros(M, N, device=dev)
out = torch.empty(M, 8, N, device=dev)
for W in sets:
    ops.linear_fwd(X, W, b, L.ACT_RELU, out=out)
torch.cuda.synchronize()
ts = []
for it in range(24):
    W = sets[it % len(sets)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.linear_fwd(X, W, b, L.ACT_RELU, out=out)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
us = ts[len(ts) // 2]
print(f"M={M} N={N} dbg={os.environ.get('PGF_LS_DBG', '0')} stages={os.environ.get('PGF_LS_STAGES', '8')}: median {us:.1f} us, "
      f"{M * N * K * 4 / us / 1e3:.0f} GB/s, min {ts[0]:.1f} us ({len(sets)} weight sets)")
