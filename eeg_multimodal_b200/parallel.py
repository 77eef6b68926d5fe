"""Multi-GPU plumbing for the two sharding modes of SURVEY.md section 8e.

* sweep (configs 3, 5): models are independent, so rank r simply owns a slice of the
  (eps x seed x variant) grid -- no data-path collective; per-model metrics are gathered on the
  host at epoch end.  The reference runs the same grid as a sequential Python loop
  (python/src/custom_models/compare_privacy_budget.py:50-62, past_acc.py:255-260).
* data-parallel (config 4): one model, the batch is split over ranks, gradients are summed with
  an all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests) before the Adam steps.

Everything here is backend-agnostic host logic (tested with gloo, world size 2, on CPU).
"""
from __future__ import annotations

import itertools
import os

import torch
import torch.distributed as dist


def sweep_grid(eps_list, n_seeds=1, base_seed=980616, variants=(None,)):
    """The experiment grid as a flat list of dicts, eps-major like the reference's loops
    (`for epsilon in epsilon_list`), then seeds, then init variants (model_dict/newfrac_* names)."""
    grid = []
    for eps, s, v in itertools.product(eps_list, range(n_seeds), variants):
        grid.append({"eps": float(eps), "seed": base_seed + s, "variant": v})
    for i, g in enumerate(grid):
        g["index"] = i
    return grid


def shard_models(n_models: int, world: int, rank: int):
    """Model index m -> rank m mod world.  Returns this rank's indices."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_models, world))


def batch_slice(global_batch: int, world: int, rank: int):
    """Contiguous row range [lo, hi) of a global batch for this rank (data-parallel mode).  `lo` is
    also the Philox row0, so the noise of a sample does not depend on the partitioning."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_distributed(backend=None):
    """env:// rendezvous from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return int(os.environ.get("RANK", "0")), world
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    kw = {}
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        kw["device_id"] = torch.device("cuda", local)
    dist.init_process_group(backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def make_allreduce_hook(group=None):
    """grad_hook for HeadEngine.train_step: sum the gradient buffer over ranks in place.  The
    engine scales dlogits by 1/global_batch, so the sum IS the gradient of the global mean loss."""
    def hook(t: torch.Tensor):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return hook


class OverlappedAllReduce:
    """grad_hook for HeadEngine.train_step in the data-parallel mode that overlaps the
    gradient exchange with the backward pass instead of running one all-reduce after it.

    The engine calls `bucket(t)` right after it has ENQUEUED the kernels that complete the gradient slice `t` (a
    contiguous 1-D view of its flat gradient buffer): first everything but fc_layers.0.weight -- final once the
    fc_layers.2 weight-gradient GEMM is queued -- then fc_layers.0.weight in `w1_chunks` row blocks, each produced by its
    own split-K GEMM launch.  Every bucket is summed over ranks on a side stream that waits for exactly those kernels,
    so its transfer runs under the GEMMs of the next bucket; `finish()` (before Adam) makes the compute stream wait for
    the last one.  Called as a plain function (`hook(t)`) it is the blocking all-reduce -- the engine uses that for
    dDP, whose 10 KB are needed by the next kernel.  Sums of fp32 in NCCL's fixed ring/tree order: the bucketing
    changes which elements travel together, not the value any element gets, so results equal the single all-reduce."""

    bucketed = True

    def __init__(self, group=None, w1_chunks=5, device=None, reserve_sms=0):
        self.group = group
        self.w1_chunks = int(w1_chunks)
        self.reserve_sms = int(reserve_sms)   # SMs the GEMM grids leave to the collective while a bucket is in flight
        self.active = dist.is_initialized() and dist.get_world_size(group) > 1
        self.cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == "cuda")
        self.comm = torch.cuda.Stream(device=device) if self.cuda else None
        self._pending = False
        self.n_buckets = 0

    def __call__(self, t: torch.Tensor):
        if self.active:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def bucket(self, t: torch.Tensor):
        self.n_buckets += 1
        if not self.active:
            return
        if self.comm is None:          # CPU / gloo (tests): no streams to overlap on
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            return
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        if self.reserve_sms and not self._pending:
            from . import _lib
            _lib.load().pgf_set_sm_reserve(self.reserve_sms)
        self._pending = True

    def finish(self):
        if self._pending:
            torch.cuda.current_stream().wait_stream(self.comm)
            self._pending = False
            if self.reserve_sms:
                from . import _lib
                _lib.load().pgf_set_sm_reserve(0)


def gather_metrics(local: dict, group=None):
    """Host-side gather of per-model metrics {model_index: value} from all ranks (epoch end)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(local)
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    merged = {}
    for d in out:
        merged.update(d)
    return dict(sorted(merged.items()))


def gather_results(local: dict, group=None):
    """Host-side gather of per-model result dicts {model_index: {...}} from all ranks.  Every rank must call it -- also
    one that owns no model (world size larger than the grid) and passes {} -- or the collective hangs."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(sorted(local.items()))
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    merged = {}
    for d in out:
        merged.update(d)
    return dict(sorted(merged.items()))


class StreamedEnsemble:
    """One GPU's share of a sweep as G independent model groups (HeadEngines), each on its own CUDA stream.

    At the reference's batch of 8 a step is ~14 kernels of 10-150 us; with few models per GPU each of them spends a
    visible part of its life ramping up and draining.  Independent groups issued on separate streams fill each other's
    ramp-up and tail.  Measured (bench.py --workload sweep48_b8 --streams 2): +1.5 % at 48 and at 128 models per GPU; at
    6 models per GPU the fork/join events and stream switches of this class cost more host time than the overlap returns
    (-5 %; a bare two-stream loop gains 4.7 % there, tools/stream_overlap_probe.py), so it is opt-in.
    The models stay exactly the models of the sweep grid: a group is a contiguous slice of the GPU's model list, every
    kernel, seed and Philox offset is what the single-engine path uses, so results are bit-identical to it.

    `train_step` / `eval_step` take the caller's batch on the caller's current stream: the group streams wait for
    whatever the caller has queued before (the batch upload), and -- unless join=False -- the caller's stream waits for
    the groups before the concatenated statistics are returned.  With join=False the per-group results are returned
    as a list and the caller joins later (`join()`), which lets consecutive steps of different groups overlap."""

    def __init__(self, engines):
        if not engines:
            raise ValueError("StreamedEnsemble needs at least one engine")
        self.engines = list(engines)
        self.streams = [torch.cuda.Stream(device=e.device) for e in self.engines]
        self.M = sum(e.M for e in self.engines)
        self._fork = torch.cuda.Event()

    def _run(self, method, blocks, labels, join, kw):
        cur = torch.cuda.current_stream()
        self._fork.record(cur)
        outs = []
        for e, s in zip(self.engines, self.streams):
            s.wait_event(self._fork)
            with torch.cuda.stream(s):
                outs.append(getattr(e, method)(blocks, labels, **kw))
        if not join:
            return outs
        self.join()
        for o in outs:
            for t in o.values():
                if torch.is_tensor(t):
                    t.record_stream(cur)          # allocated on a group stream, consumed on the caller's
        return {k: torch.cat([o[k] for o in outs]) for k in outs[0] if torch.is_tensor(outs[0][k])}

    def train_step(self, blocks, labels, join=True, **kw):
        return self._run("train_step", blocks, labels, join, kw)

    def eval_step(self, blocks, labels, join=True, **kw):
        return self._run("eval_step", blocks, labels, join, kw)

    def join(self):
        cur = torch.cuda.current_stream()
        for s in self.streams:
            cur.wait_stream(s)
