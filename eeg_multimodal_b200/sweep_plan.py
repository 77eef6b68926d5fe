"""The reference step at the reference batch size as one constant, graph-replayed launch sequence.

`SweepStepPlan` binds a `HeadEngine` (fp32 path, the models of one GPU's share of the sweep) to a dataset that is
RESIDENT in HBM and drives `pgf_sweep_plan_*` (include/pgfuse.h, csrc/sweep_step.cu): per training step
(past_acc.py:198-212) the host issues one `cudaGraphLaunch` -- or one per `steps_per_graph` steps -- and nothing else.
The batch of a step is rows `src_rows[cursor : cursor+B]` of the resident blocks (a shuffled epoch = one permutation
uploaded per epoch, data.py:37-45); the cursor, the Philox offsets and the Adam step counts / bias corrections live in
a 64-byte device struct that the step's last kernel advances.

Results are bit-identical to `HeadEngine.train_step` on the same batches (tests/test_gpu_sweep_step.py): the plan
runs the same kernels with the same arithmetic, it only removes launches and host work.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from . import ops


class SweepStepPlan:
    def __init__(self, eng, blocks, labels, batch_size, src_rows=None, dp_pass=True, use_pdl=True, wrap=True):
        if eng.precision != "fp32":
            raise ValueError("the fused sweep step is the fp32 reference-batch path (precision='fp32')")
        if not 1 <= int(batch_size) <= 8:
            raise ValueError("the fused sweep step handles the reference batch sizes 1..8; larger batches use HeadEngine.train_step")
        if eng.noise != "philox":
            raise ValueError("the fused sweep step draws its noise in-kernel (noise='philox')")
        self.eng, self.B = eng, int(batch_size)
        dev = eng.device
        blocks = [ops._chk(b, torch.float32, "dataset block") for b in blocks]
        if len(blocks) != len(eng.dims) or any(b.dim() != 2 or b.shape[1] != d for b, d in zip(blocks, eng.dims)):
            raise ValueError(f"dataset blocks must be [N, d_i] with widths {eng.dims}")
        n = blocks[0].shape[0]
        labels = ops._chk(labels.reshape(-1), torch.int64, "labels").contiguous()
        if any(b.shape[0] != n for b in blocks) or labels.shape[0] != n:
            raise ValueError("all dataset blocks and the labels must have the same number of rows")
        if n < self.B:
            raise ValueError("the resident dataset holds fewer rows than one batch")
        self.blocks, self.labels, self.n = blocks, labels, n
        self.src_rows = None
        self.dp_pass = bool(dp_pass)
        self.cursor = 0
        M, D, H = eng.M, eng.D, eng.H
        self.state = torch.zeros(8, dtype=torch.int64, device=dev)           # pgf_step_state, 64 bytes
        nbytes = L.query("pgf_sweep_plan_workspace", M, self.B, D, H)
        self.ws = torch.empty((nbytes + 255) // 256 * 256 + 256, dtype=torch.uint8, device=dev)
        ws_ptr = (self.ws.data_ptr() + 255) // 256 * 256
        self.stats_dp = eng._buf("stats_dp", (M, 4), torch.float32)
        self.stats_model = eng._buf("stats_model", (M, 4), torch.float32)
        self.logits = eng._buf("logits_model", (M, self.B, 2), torch.float32)
        self.pred = eng._buf("pred_model", (M, self.B), torch.int64)
        self.coef = eng._ensure_coef()
        d = L.SweepDesc()
        d.n_models, d.B, d.H = M, self.B, H
        dims = list(eng.dims) + [0] * (3 - len(eng.dims))
        d.d0, d.d1, d.d2 = dims
        d.dp_pass, d.fixed_formula, d.use_pdl = int(self.dp_pass), int(eng.fixed), int(bool(use_pdl))
        d.tau, d.lr, d.beta1, d.beta2, d.adam_eps = eng.tau, eng.lr, eng.betas[0], eng.betas[1], eng.adam_eps
        for i, b in enumerate(blocks):
            setattr(d, f"x{i}", b.data_ptr())
            setattr(d, f"ld{i}", b.stride(0))
        d.labels = labels.data_ptr()
        d.src_rows = None
        d.n_rows = n if wrap else 0
        d.params, d.adam_m, d.adam_v, d.grads, d.P = eng.flat.data_ptr(), eng.m.data_ptr(), eng.v.data_ptr(), eng.grad.data_ptr(), eng.P
        for name in ("W1", "b1", "W2", "b2", "Wc", "bc"):
            setattr(d, "off_" + name, eng.layout[name][0])
        d.DP, d.DP_m, d.DP_v, d.dDP = eng.DP.data_ptr(), eng.DP_m.data_ptr(), eng.DP_v.data_ptr(), eng.dDP.data_ptr()
        d.coef, d.exp_eps, d.seeds, d.row0 = self.coef.data_ptr(), eng.exp_eps_dev.data_ptr(), eng.seeds_dev.data_ptr(), 0
        d.stats_dp, d.stats_model = self.stats_dp.data_ptr(), self.stats_model.data_ptr()
        d.logits, d.pred, d.state = self.logits.data_ptr(), self.pred.data_ptr(), self.state.data_ptr()
        d.workspace, d.workspace_bytes = ws_ptr, nbytes
        self._desc = d
        self._plan = None
        self._captured = 0
        self._stream = None
        self._create()
        self.launches_per_step = int(L.load().pgf_sweep_plan_launches_per_step(self._plan))

    # ---- plan object -----------------------------------------------------------------------------
    def _create(self):
        if self._plan is not None:
            L.call("pgf_sweep_plan_destroy", self._plan)
        h = C.c_void_p()
        L.call("pgf_sweep_plan_create", C.byref(self._desc), C.byref(h))
        self._plan = h
        self._captured = 0
        self._needs_reset = True

    def close(self):
        if self._plan is not None:
            L.call("pgf_sweep_plan_destroy", self._plan)
            self._plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- epoch / state ---------------------------------------------------------------------------
    def set_rows(self, src_rows, cursor=0):
        """Row order of the coming epoch: device int64 tensor (a permutation of the resident rows, any length >= one
        batch) or None for the dataset's own order.  A change of the tensor's address re-creates the plan."""
        ptr = None
        if src_rows is not None:
            src_rows = ops._chk(src_rows, torch.int64, "src_rows").contiguous()
            ptr = src_rows.data_ptr()
        old = None if self.src_rows is None else self.src_rows.data_ptr()
        self.src_rows = src_rows
        self.cursor = int(cursor)
        if ptr != old:
            self._desc.src_rows = ptr
            self._desc.n_rows = (src_rows.numel() if src_rows is not None else self.n) if self._desc.n_rows else 0
            cap = self._captured
            self._create()
            if cap:
                self._recapture = cap
        self.sync_state()

    def sync_state(self):
        """Write the engine's step counters and this plan's cursor into the device state."""
        e = self.eng
        L.call("pgf_step_state_set", self.state.data_ptr(), e.noise_offset, e.t_dp, e.t_model, self.cursor, e.lr, e.betas[0], e.betas[1],
               ops._stream())
        self._synced = (e.noise_offset, e.t_dp, e.t_model)

    def capture(self, steps_per_graph=1):
        """Record `steps_per_graph` steps into a CUDA graph on the current stream (after at least one direct step, so that
        every lazy per-kernel attribute is set)."""
        if self._needs_reset:
            L.call("pgf_sweep_plan_reset", self._plan, ops._stream())
            self._needs_reset = False
        L.call("pgf_sweep_plan_capture", self._plan, ops._stream(), int(steps_per_graph))
        self._captured = int(steps_per_graph)
        self._stream = ops._stream()

    # ---- stepping --------------------------------------------------------------------------------
    def run(self, n_steps=1):
        """Enqueue n_steps training steps on the current stream.  Returns the LIVE per-model statistics buffer of pass 2
        ([M,4] = loss, n_correct, accuracy, B of the last step): it is overwritten by the next step, copy what you keep."""
        e = self.eng
        if getattr(self, "_synced", None) != (e.noise_offset, e.t_dp, e.t_model):
            self.sync_state()              # the engine stepped outside the plan (train_step / eval_step) since the last run
        if e._coef_key != e._coef_state():
            e._ensure_coef()
        stream = ops._stream()
        if self._needs_reset:
            L.call("pgf_sweep_plan_reset", self._plan, stream)
            self._needs_reset = False
        if getattr(self, "_recapture", 0):
            L.call("pgf_sweep_plan_capture", self._plan, stream, self._recapture)
            self._captured, self._stream, self._recapture = self._recapture, stream, 0
        L.call("pgf_sweep_plan_run", self._plan, stream, int(n_steps))
        L.launch_count += n_steps * self.launches_per_step - 1       # L.call counted one launch for the entry point
        per = 2 if self.dp_pass else 1
        e.noise_offset += per * n_steps
        e.t_dp += n_steps if self.dp_pass else 0
        e.t_model += n_steps
        e._coef_key = e._coef_state()      # the DP pass refreshed the rows in place
        self._synced = (e.noise_offset, e.t_dp, e.t_model)
        for _ in range(n_steps):           # host mirror of step_advance_kernel's cursor rule
            c = self.cursor + self.B
            if self._desc.n_rows and c + self.B > self._desc.n_rows:
                c = 0
            self.cursor = c
        return self.stats_model

    def result(self):
        st = self.stats_model
        return dict(loss=st[:, 0], acc=st[:, 2], n_correct=st[:, 1], stats=st, pred=self.pred, logits=self.logits)
