"""CPU, world size 2 over gloo: the host-side multi-GPU logic.

* data-parallel mode: each rank differentiates its contiguous slice of the global batch with the
  loss scaled by 1/global_batch (what HeadEngine does through `grad_scale`), the all-reduce hook
  sums the gradient buffers, and every rank ends up with the full-batch gradient of the mean loss.
* sweep mode: ranks own disjoint model subsets and only gather per-model metrics at epoch end.
The GPU kernels are not involved (the oracle supplies the arithmetic); NCCL replaces gloo on the box.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from eeg_multimodal_b200 import parallel
from oracle import head_oracle as ho

D, H, GB = 64, 32, 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _grads(blocks, label, p, lap, gum, scale):
    q = p.clone(requires_grad=True)
    pred = ho.head_forward(blocks, q, 1.0, lap, gum, hard=True)
    loss = F.cross_entropy(pred, label.squeeze(1), reduction="sum") * scale
    loss.backward()
    return torch.cat([t.grad.reshape(-1) for t in q.tensors()])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, w = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)                       # same data and noise on every rank
    blocks = [torch.rand(GB, D // 2), torch.rand(GB, D // 2)]
    label = (torch.rand(GB, 1) < 0.66).long()
    lap, gum = ho.replay_reference_draws(3, GB, D)
    p = ho.make_params(D, H, seed=1)
    full = _grads(blocks, label, p, lap, gum, 1.0 / GB)
    lo, hi = parallel.batch_slice(GB, world, rank)
    mine = _grads([b[lo:hi] for b in blocks], label[lo:hi], p, lap[lo:hi], gum[:, lo:hi], 1.0 / GB)
    hook = parallel.make_allreduce_hook()
    bucketed = mine.clone()
    hook(mine)
    err = float((mine - full).abs().max() / full.abs().max())
    # the same exchange in buckets (what the data-parallel GPU step does while the backward pass is still running)
    ov = parallel.OverlappedAllReduce(w1_chunks=3)
    k = bucketed.numel() // 3
    for a, b in ((2 * k, bucketed.numel()), (0, k), (k, 2 * k)):
        ov.bucket(bucketed[a:b])
    ov.finish()
    assert ov.active and ov.n_buckets == 3 and torch.equal(bucketed, mine)
    # shared sweep batch: every rank uploads its rows only, the exchange fills in the peers' rows
    fan = parallel.SharedBatchFanout(GB, "cpu")
    shared = torch.arange(GB * 3, dtype=torch.float32).view(GB, 3)
    lab = torch.arange(GB)
    dev_t, dev_l = torch.full((GB, 3), -1.0), torch.full((GB,), -1, dtype=torch.int64)
    assert fan.register({0: [dev_t, dev_l]}) == "nccl"          # CPU / gloo: the collective transport
    fan.upload(0, [fan.host_slice(shared), fan.host_slice(lab)])
    assert (fan.lo, fan.hi) == parallel.batch_slice(GB, world, rank) and torch.equal(dev_t, shared) and torch.equal(dev_l, lab)
    fan.close()
    # sweep mode: disjoint ownership, metrics gathered on every rank
    own = parallel.shard_models(5, world, rank)
    merged = parallel.gather_metrics({m: 0.9 + 0.01 * m for m in own})
    torch.save({"err": err, "own": own, "merged": merged}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world2_gloo_allreduce_and_gather(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    for r in res:
        assert r["err"] < 1e-5                                   # summed slice gradients == full-batch gradient
        assert list(r["merged"].keys()) == [0, 1, 2, 3, 4]          # every model reported exactly once
    assert res[0]["own"] == [0, 2, 4] and res[1]["own"] == [1, 3]
    assert res[0]["merged"] == res[1]["merged"]


def _worker_small_grid(rank, world, port, out_dir):
    """A 1-model grid on 2 ranks: rank 1 owns nothing and must still take part in every collective of the driver's tail
    (eeg_multimodal_b200/train.py: gather_metrics + gather_results), or rank 0 hangs before results.pth is written."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    parallel.init_distributed("gloo")
    grid = parallel.sweep_grid([1.0], 1)
    mine = [grid[i] for i in parallel.shard_models(len(grid), world, rank)]
    local = {g["index"]: {"best_acc": 0.75, "reference": {"Accuracy": torch.tensor([0.75])}} for g in mine}
    merged = parallel.gather_metrics({i: r["best_acc"] for i, r in local.items()})
    everything = parallel.gather_results(local)
    torch.save({"n_mine": len(mine), "merged": merged, "keys": list(everything)}, os.path.join(out_dir, f"s{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world2_gloo_rank_without_models_joins_the_gathers(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker_small_grid, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"s{r}.pt") for r in range(world)]
    assert [r["n_mine"] for r in res] == [1, 0]
    for r in res:
        assert r["merged"] == {0: 0.75} and r["keys"] == [0]


def test_single_process_paths_need_no_process_group():
    os.environ.pop("WORLD_SIZE", None)
    assert parallel.init_distributed("gloo") == (int(os.environ.get("RANK", "0")), 1)
    t = torch.ones(3)
    parallel.make_allreduce_hook()(t)                               # no-op without a group
    assert torch.equal(t, torch.ones(3))
    assert parallel.gather_metrics({1: 0.5}) == {1: 0.5}
    assert parallel.gather_results({2: {"a": 1}, 1: {"a": 0}}) == {1: {"a": 0}, 2: {"a": 1}}
