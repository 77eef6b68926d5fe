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


class SharedBatchFanout:
    """The batch of a sweep step crosses the host link ONCE per node instead of once per GPU.

    Every model of the eps x seed sweep reads the same dataset in the same order (the reference runs all of them with one
    seed over one DataLoader, compare_privacy_budget.py:50-62), so with the models sharded over the GPUs of a box every
    rank used to upload the same batch: eight concurrent 671 MB copies per step saturate the host side and cost 12 % of the
    8-GPU end-to-end rate.  Here rank r uploads rows [r*B/N, (r+1)*B/N) from its pinned host slice and the slices are
    exchanged over NVLink, under the previous step's GEMMs.  Two transports:

    mode='nccl'  (default) in-place NCCL all-gather on a communicator of its own limited to `max_ctas` CTAs, with the GEMM
                 grids leaving that many SMs free (pgf_set_sm_reserve): a collective CTA that queues behind a persistent
                 grid, or a GEMM cluster that queues behind a collective CTA, costs far more than the SMs given up.
                 Measured at 2 GPUs (profiles/r2_fanout.txt): an upload takes 8.1 ms with 8 CTAs (6.1 ms of it the H2D),
                 13.8 ms with 2; end to end 0.88 / 0.94 / 0.95 / 0.96 of the HBM-resident rate with 1 / 2 / 4 / 8 CTAs
                 against 0.98 for plain per-GPU uploads -- the exchange only pays where the host link is the limit
                 (8 GPUs: 0.88 without it).
    mode='p2p'   the destination buffers of every rank are mapped into every process (CUDA IPC) and each rank PUSHES its
                 rows into its peers' buffers with cudaMemcpyPeerAsync (pgf_memcpy_peer_async); two scalar all-reduces
                 per upload order the pushes against the consumers on the other ranks (buffer free / rows landed).  On
                 the virtualised boxes of this pool copies into another process's IPC mapping are staged through the
                 host (measured 25 GB/s and 17 ms of blocked host time per upload), so it is not the default.
    `register` falls back from 'p2p' to 'nccl' on every rank if any rank cannot map a peer buffer."""

    def __init__(self, batch, device, mode="nccl", max_ctas=8, reserve_sms=None):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if batch % self.world:
            raise ValueError(f"the shared batch of {batch} rows does not split evenly over {self.world} ranks")
        self.rows = batch // self.world
        self.lo, self.hi = self.rank * self.rows, (self.rank + 1) * self.rows
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.mode = mode if (self.cuda and self.world > 1) else "nccl"
        self.max_ctas = int(max_ctas) if self.mode == "nccl" else 1     # p2p: the communicator only carries the two scalar barriers
        # SMs the GEMM grids leave free while this object is live: the collective's CTAs ('nccl'), or one CTA pair for the
        # barrier kernels ('p2p' -- without it a barrier waits for a GEMM boundary on every rank, and while it spins for
        # the slowest rank it holds an SM that the next GEMM's static tile schedule counts on: measured 0.91 instead of 0.97)
        self.reserve_sms = (self.max_ctas if self.mode == "nccl" else 2) if reserve_sms is None else int(reserve_sms)
        self.group = None
        self.peers = {}          # slot -> list over ranks of the list of that rank's destination tensors (None for self)
        self._flag = None
        self._reserved = False

    # ---- setup ----------------------------------------------------------------------------------
    def _nccl_group(self):
        if self.group is None and self.world > 1:
            kw = {}
            if self.cuda:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = self.max_ctas
                opts.config.min_ctas = 1
                kw = dict(backend="nccl", pg_options=opts)
            self.group = dist.new_group(**kw)        # collective: every rank gets here together (see register)
            if self.cuda and self.reserve_sms > 0:
                from . import _lib
                _lib.load().pgf_set_sm_reserve(self.reserve_sms)
                self._reserved = True
        return self.group

    def register(self, slots):
        """slots: {slot: [destination tensors of that slot, each [B, ...]]} -- the buffers `upload` will fill.  Called once
        by every rank with the same structure.  In 'p2p' mode the buffers' CUDA IPC handles are exchanged and mapped."""
        self.slots = {k: list(v) for k, v in slots.items()}
        if self.world == 1:
            return self.mode
        if self.mode == "p2p":
            try:
                mine = {k: [(t.untyped_storage()._share_cuda_(), t.storage_offset(), tuple(t.shape), tuple(t.stride()), t.dtype)
                            for t in v] for k, v in self.slots.items()}
                ok = True
            except Exception as e:          # noqa: BLE001 -- any failure means "use the other transport", on every rank
                mine, ok = repr(e), False
            everyone = [None] * self.world
            dist.all_gather_object(everyone, (ok, mine))
            ok = all(o for o, _ in everyone)
            if ok:
                try:
                    for k in self.slots:
                        per_rank = []
                        for r, (_, handles) in enumerate(everyone):
                            if r == self.rank:
                                per_rank.append(None)
                                continue
                            ts = []
                            for h, off, shape, stride, dtype in handles[k]:
                                st = torch.UntypedStorage._new_shared_cuda(*h)
                                ts.append(torch.empty(0, dtype=dtype, device=st.device).set_(st, off, shape, stride))
                            per_rank.append(ts)
                        self.peers[k] = per_rank
                    self._flag = torch.zeros(1, device=self.device)
                except Exception:           # noqa: BLE001
                    ok = False
            agreed = [None] * self.world
            dist.all_gather_object(agreed, ok)
            if not all(agreed):
                self.peers, self.mode = {}, "nccl"
        self._nccl_group()
        return self.mode

    def host_slice(self, t: torch.Tensor) -> torch.Tensor:
        """This rank's rows of a host-side batch tensor."""
        return t[self.lo:self.hi]

    # ---- per step -------------------------------------------------------------------------------
    def upload(self, slot, host_slices):
        """host_slices[i] ([B/N, ...], pinned) -> rows [lo, hi) of the slot's i-th destination tensor on EVERY rank.
        Enqueued on the current stream, which must already wait for this rank's consumers of the slot; returns once the
        work is queued.  After it (in stream order) the slot holds the whole batch."""
        dst = self.slots[slot]
        if self.world == 1:
            for hs, dt in zip(host_slices, dst):
                dt.copy_(hs, non_blocking=True)
            return
        if self.mode == "nccl":
            for hs, dt in zip(host_slices, dst):
                mine = dt[self.lo:self.hi]
                mine.copy_(hs, non_blocking=True)
                dist.all_gather_into_tensor(dt, mine, group=self._nccl_group())
            return
        dist.all_reduce(self._flag, group=self.group)     # every rank's consumers have released the slot
        for hs, dt in zip(host_slices, dst):
            dt[self.lo:self.hi].copy_(hs, non_blocking=True)
        from . import _lib
        stream = torch.cuda.current_stream().cuda_stream
        for r in range(1, self.world):                    # staggered so that the ranks do not all push to the same peer at once
            q = (self.rank + r) % self.world
            for dt, pt in zip(dst, self.peers[slot][q]):
                src, to = dt[self.lo:self.hi], pt[self.lo:self.hi]
                # raw cudaMemcpyPeerAsync on this rank's stream (a torch copy_ between devices of two processes is staged
                # and blocks the host: measured 17 ms per upload)
                _lib.call("pgf_memcpy_peer_async", to.data_ptr(), pt.device.index, src.data_ptr(), dt.device.index,
                          src.numel() * src.element_size(), stream)
        dist.all_reduce(self._flag, group=self.group)     # every rank's pushes have landed

    def close(self):
        if self._reserved:
            from . import _lib
            _lib.load().pgf_set_sm_reserve(0)
            self._reserved = False
        self.peers = {}


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
