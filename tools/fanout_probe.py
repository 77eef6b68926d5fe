"""Raw cost of one shared-batch upload (no compute running): H2D of 1/N of the batch, then the exchange over NVLink.
torchrun --nproc-per-node N tools/fanout_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eeg_multimodal_b200 import parallel  # noqa: E402

rank, world = parallel.init_distributed()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
B, dims = 65536, (2048, 512)


def run(mode, ctas=8, K=10):
    fan = parallel.SharedBatchFanout(B, dev, mode=mode, max_ctas=ctas, reserve_sms=0)
    host = [[torch.rand(B // world, d).pin_memory() for d in dims] + [torch.zeros(B // world, dtype=torch.int64).pin_memory()] for _ in range(2)]
    devb = {s: [torch.empty(B, d, device=dev) for d in dims] + [torch.empty(B, dtype=torch.int64, device=dev)] for s in range(2)}
    got = fan.register(devb)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(3):
            fan.upload(i % 2, host[i % 2])
        st.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fan.upload(i % 2, host[i % 2])
        e1.record()
        t_host = time.perf_counter() - t0
        st.synchronize()
    ok = all(torch.equal(devb[0][j][fan.lo:fan.hi].cpu(), host[(K - 2) % 2 if False else 0][j]) for j in range(3)) if K % 2 == 0 else None
    if rank == 0:
        print(f"mode={got} ctas={ctas}: {e0.elapsed_time(e1) / K:.2f} ms per upload on the device, {t_host / K * 1e3:.2f} ms host enqueue; own rows intact: {ok}", flush=True)
    fan.close()
    dist.barrier()


# plain H2D of the slice alone
host = [torch.rand(B // world, d).pin_memory() for d in dims]
d = [torch.empty(B, dd, device=dev) for dd in dims]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    for h, t in zip(host, d):
        t[: B // world].copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"H2D of 1/{world} of the batch alone: {e0.elapsed_time(e1) / 10:.2f} ms", flush=True)
run("nccl", 8)
run("nccl", 2)
run("p2p")
dist.destroy_process_group()
