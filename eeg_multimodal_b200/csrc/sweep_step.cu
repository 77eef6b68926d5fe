// The reference's training step at its own batch size (B <= 8) for a whole eps x seed sweep, as ONE constant launch
// sequence: device-resident step state, programmatic dependent launch along the chain, CUDA-graph replay.
//
// Reference step replaced (past_acc.py:198-212 == base_train.py:183-210), for every model of the sweep at once:
//   pass 1  hard=False forward -> cal_loss -> backward -> DP_optimizer.step()        (skipped when dp_pass == 0: train.py:100-105)
//   pass 2  hard=True  forward -> cal_loss -> backward -> model_optimizer.step()
// and the DataLoader(shuffle=True) batch fetch in front of it (data.py:37-45): the batch is rows
// src_rows[cursor .. cursor+B) of a dataset resident in HBM, gathered inside the first kernel.
//
// Why a dedicated path: at B=8 a layer is a weight-streaming pass of 20-150 us and the step is a chain of ~20
// dependent launches; issued one by one from the host every launch costs >= 6.6 us however little it does (r1: 0.44 ms
// per step of 6 models against 0.24 ms of HBM time).  Here
//   * everything that changes from step to step (Philox offsets, Adam step counts and bias corrections, the batch
//     cursor) lives in a 64-byte device struct (StepState) that the kernels add to their constant arguments, so the
//     step is capturable once and replayable forever;
//   * the chain is 13 launches instead of 21: the slab reductions of the dX kernels run in the last CTA to finish
//     (no finalize launches), Adam(DP) + the coefficient refresh ride in the dDP kernel, Adam(classifier) in the
//     pass-2 loss kernel, both fc layers' gradient+Adam are one launch, whose last CTA also advances the step state;
//   * every kernel is launched as a programmatic dependent of its predecessor: its CTAs are resident and have their
//     first weight vectors in flight when the predecessor's last CTA retires.
// Same kernels, same summation orders and the same Adam arithmetic as the one-call-per-kernel path: results are
// bit-identical to it (tests/test_gpu_sweep_step.py).
#include <new>

#include "../../include/pgfuse.h"
#include "pgf_kernels.cuh"

namespace pgf {

__global__ void step_state_set_kernel(StepState* st, long long noise_offset, long long t_dp, long long t_model, long long cursor,
                                      float lr, float b1, float b2) {
  st->noise_offset = noise_offset;
  st->t_dp = t_dp;
  st->t_model = t_model;
  st->cursor = cursor;
  st->lr = lr; st->b1 = b1; st->b2 = b2; st->pad = 0.f;
  const AdamCoef cd = make_adam_coef(t_dp + 1, lr, b1, b2, 0.f, 1.f);
  const AdamCoef cm = make_adam_coef(t_model + 1, lr, b1, b2, 0.f, 1.f);
  st->dp_step_size = cd.step_size; st->dp_bc2_sqrt = cd.bc2_sqrt;
  st->model_step_size = cm.step_size; st->model_bc2_sqrt = cm.bc2_sqrt;
}


static inline size_t align256(size_t n) { return (n + 255) & ~static_cast<size_t>(255); }

struct SweepPlan {
  pgf_sweep_desc d;
  int D;
  // workspace carve-up
  float *X, *H1, *H2, *dZ2, *dZ1, *dX, *dx_part, *ce_ws;
  unsigned int* counters;
  size_t dx_part_bytes, ce_ws_bytes;
  bool pdl;
  cudaGraphExec_t exec;
  int graph_steps;
  cudaStream_t graph_stream;

  static size_t workspace_bytes(int M, int B, int D, int H, size_t* offs) {
    const size_t act_d = align256(static_cast<size_t>(M) * B * D * sizeof(float));
    const size_t act_h = align256(static_cast<size_t>(M) * B * H * sizeof(float));
    size_t dxp = linear_dx_workspace(B, H, D, M);
    const size_t dxp1 = linear_dx_workspace(B, D, D, M);
    if (dxp1 > dxp) dxp = dxp1;
    if (M >= PGF_WIDE_MODELS) {   // many models per GPU: the slab kernels of linear_wide.cu take dX and gradient+Adam
      const size_t w0 = linear_dx_workspace_wide(B, H, D, M), w1 = linear_dx_workspace_wide(B, D, D, M);
      if (w0 > dxp) dxp = w0;
      if (w1 > dxp) dxp = w1;
    }
    dxp = align256(dxp);
    const size_t ce = align256(cls_ce_workspace(B, H, M));
    const size_t cnt = align256((static_cast<size_t>(linear_dx_counters(B, D, M)) + 1) * sizeof(unsigned int));   // + the step-advance ticket
    size_t o = 0;
    size_t tmp[10];
    size_t* p = offs ? offs : tmp;
    p[0] = o; o += act_d;   // X
    p[1] = o; o += act_d;   // H1
    p[2] = o; o += act_h;   // H2
    p[3] = o; o += act_h;   // dZ2
    p[4] = o; o += act_d;   // dZ1
    p[5] = o; o += act_d;   // dX
    p[6] = o; o += dxp;     // dx slab partials
    p[7] = o; o += ce;      // cls_ce workspace
    p[8] = o; o += cnt;     // dx counters
    p[9] = dxp;
    return o;
  }

  int pass(cudaStream_t s, bool model_pass, bool first_pdl, unsigned int offset_delta);
  int enqueue(cudaStream_t s, bool first_pdl);
};

int SweepPlan::pass(cudaStream_t s, bool model_pass, bool first_pdl, unsigned int offset_delta) {
  const int M = d.n_models, B = d.B, H = d.H;
  const StepState* st = static_cast<const StepState*>(d.state);
  float* coef_w = d.coef;
  float* coef_eh = d.coef + static_cast<long long>(M) * D;
  float* coef_de = d.coef + 2LL * M * D;
  int rc;
  {  // (a1-a7) gather + normalise + perturb, all models
    PdlScope scope(pdl && first_pdl);
    PerturbFwdArgs a = {};
    a.x[0] = d.x0; a.x[1] = d.x1; a.x[2] = d.x2;
    a.ld[0] = d.ld0; a.ld[1] = d.ld1; a.ld[2] = d.ld2;
    a.d[0] = d.d0; a.d[1] = d.d1; a.d[2] = d.d2;
    a.D = D; a.B = B; a.n_models = M; a.s_coef = D; a.s_out = static_cast<long long>(B) * D;
    a.model_seeds = d.seeds;
    a.w = coef_w; a.eps_hat = coef_eh;
    a.offset = offset_delta; a.row0 = d.row0;
    a.tau = d.tau; a.inv_tau = 1.0f / d.tau; a.hard = model_pass ? 1 : 0;
    a.out = X; a.ld_out = D;
    a.st = st; a.src_rows = d.src_rows; a.gather = 1; a.n_rep = 1;
    rc = perturb_gate_fwd(a, PGF_NOISE_PHILOX, PGF_DT_F32, false, s);
    if (rc != PGF_OK) return rc;
  }
  PdlScope scope(pdl);
  float* W1 = d.params + d.off_W1; float* b1 = d.params + d.off_b1;
  float* W2 = d.params + d.off_W2; float* b2 = d.params + d.off_b2;
  float* Wc = d.params + d.off_Wc; float* bc = d.params + d.off_bc;
  {  // (a8) fc_layers
    LinFwdArgs a;
    a.X = X; a.ldx = D; a.sX = static_cast<long long>(B) * D; a.W = W1; a.sW = d.P; a.bias = b1; a.sb = d.P;
    a.Y = H1; a.ldy = D; a.sY = static_cast<long long>(B) * D; a.B = B; a.N = D; a.K = D; a.act = PGF_ACT_RELU;
    rc = linear_fwd(a, M, s);
    if (rc != PGF_OK) return rc;
    a.X = H1; a.W = W2; a.bias = b2; a.Y = H2; a.ldy = H; a.sY = static_cast<long long>(B) * H; a.N = H; a.act = PGF_ACT_TANH;
    rc = linear_fwd(a, M, s);
    if (rc != PGF_OK) return rc;
  }
  {  // (a9-a11) classifier + loss + dZ2 (+ pass 2: dWc, dbc and their Adam update)
    CeArgs a = {};
    a.h = H2; a.ldh = H; a.sh = static_cast<long long>(B) * H; a.Wc = Wc; a.sWc = d.P; a.bc = bc; a.sbc = d.P;
    a.labels = d.labels; a.slab = 0;
    a.logits = model_pass ? d.logits : nullptr; a.slogits = static_cast<long long>(B) * 2;
    a.pred = model_pass ? d.pred : nullptr; a.spred = B;
    a.dz = dZ2; a.lddz = H; a.sdz = static_cast<long long>(B) * H;
    a.B = B; a.H = H; a.grad_scale = 1.0f / static_cast<float>(B); a.through_tanh = 1;
    a.st = st; a.src_rows = d.src_rows; a.gather = 1;
    float* gWc = nullptr; float* gbc = nullptr;
    if (model_pass) {
      gWc = d.grads + d.off_Wc; gbc = d.grads + d.off_bc;
      a.adam_m = d.adam_m + d.off_Wc; a.adam_v = d.adam_v + d.off_Wc;
      a.adam_mb = d.adam_m + d.off_bc; a.adam_vb = d.adam_v + d.off_bc;
      a.adam_c = make_adam_coef(1, d.lr, d.beta1, d.beta2, d.adam_eps, 1.f);   // step-dependent fields come from the state
    }
    rc = cls_ce(a, PGF_DT_F32, PGF_DT_F32, 1, M, 1.0f / static_cast<float>(B), model_pass ? d.stats_model : d.stats_dp, gWc, d.P,
                gbc, d.P, nullptr, 0, ce_ws, ce_ws_bytes, s);
    if (rc != PGF_OK) return rc;
  }
  // (a11) dZ1 = (dZ2 . W2) * relu'(H1)
  const bool wide = M >= PGF_WIDE_MODELS;   // millisecond launches: page locality decides, not ramp-up and tail (linear_wide.cu)
  rc = wide ? linear_bwd_dx_wide(dZ2, H, static_cast<long long>(B) * H, W2, d.P, H1, PGF_ACT_RELU, D, static_cast<long long>(B) * D,
                                 dZ1, D, static_cast<long long>(B) * D, B, H, D, M, dx_part, dx_part_bytes, s)
            : linear_bwd_dx(dZ2, H, static_cast<long long>(B) * H, W2, d.P, H1, PGF_ACT_RELU, D, static_cast<long long>(B) * D, dZ1, D,
                            static_cast<long long>(B) * D, B, H, D, M, dx_part, dx_part_bytes, s, counters);
  if (rc != PGF_OK) return rc;
  if (!model_pass) {
    // dX = dZ1 . W1, then dDP + Adam(DP) + coefficient refresh
    rc = wide ? linear_bwd_dx_wide(dZ1, D, static_cast<long long>(B) * D, W1, d.P, nullptr, PGF_ACT_RELU, 0, 0, dX, D,
                                   static_cast<long long>(B) * D, B, D, D, M, dx_part, dx_part_bytes, s)
              : linear_bwd_dx(dZ1, D, static_cast<long long>(B) * D, W1, d.P, nullptr, PGF_ACT_RELU, 0, 0, dX, D,
                              static_cast<long long>(B) * D, B, D, D, M, dx_part, dx_part_bytes, s, counters);
    if (rc != PGF_OK) return rc;
    DpAdamFuse f;
    f.DP = d.DP; f.DP_m = d.DP_m; f.DP_v = d.DP_v;
    f.c = make_adam_coef(1, d.lr, d.beta1, d.beta2, d.adam_eps, 1.f);
    f.exp_eps = d.exp_eps; f.fixed = d.fixed_formula;
    f.w = coef_w; f.eps_hat = coef_eh; f.deps = coef_de;
    return perturb_gate_bwd_dp(dX, PGF_DT_F32, D, static_cast<long long>(B) * D, B, D, M, PGF_NOISE_PHILOX, nullptr, 0, 0, d.seeds,
                               offset_delta, d.row0, coef_de, D, dx_part, dx_part_bytes, d.dDP, D, 0, s, st, &f);
  }
  // (a11,a12) weight gradients recomputed inside Adam, both fc layers in one launch (fc_layers.2 first, as the
  // one-call-per-kernel path orders them)
  LinAdamArgs a = {};
  a.n_layers = 2; a.sP = d.P; a.B = B; a.st = st;
  a.c = make_adam_coef(1, d.lr, d.beta1, d.beta2, d.adam_eps, 1.f);
  LinAdamLayer& l2 = a.l[0];
  l2.dY = dZ2; l2.ldy = H; l2.sdY = static_cast<long long>(B) * H; l2.X = H1; l2.ldx = D; l2.sX = static_cast<long long>(B) * D;
  l2.W = W2; l2.mW = d.adam_m + d.off_W2; l2.vW = d.adam_v + d.off_W2;
  l2.bias = b2; l2.mb = d.adam_m + d.off_b2; l2.vb = d.adam_v + d.off_b2; l2.N = H; l2.K = D;
  LinAdamLayer& l1 = a.l[1];
  l1.dY = dZ1; l1.ldy = D; l1.sdY = static_cast<long long>(B) * D; l1.X = X; l1.ldx = D; l1.sX = static_cast<long long>(B) * D;
  l1.W = W1; l1.mW = d.adam_m + d.off_W1; l1.vW = d.adam_v + d.off_W1;
  l1.bias = b1; l1.mb = d.adam_m + d.off_b1; l1.vb = d.adam_v + d.off_b1; l1.N = D; l1.K = D;
  // the step's last kernel also advances the device state (Philox offsets, Adam step counts + bias corrections, cursor)
  a.adv.st = static_cast<StepState*>(d.state);
  a.adv.counter = counters + linear_dx_counters(B, D, M);
  a.adv.d_noise = d.dp_pass ? 2 : 1; a.adv.d_tdp = d.dp_pass ? 1 : 0; a.adv.d_tmodel = 1;
  a.adv.d_cursor = B; a.adv.n_rows = d.n_rows;
  return wide ? linear_adam_step_wide(a, M, s) : linear_adam_step(a, M, s);
}

int SweepPlan::enqueue(cudaStream_t s, bool first_pdl) {
  int rc;
  if (d.dp_pass) {
    rc = pass(s, false, first_pdl, 0u);
    if (rc != PGF_OK) return rc;
  }
  return pass(s, true, d.dp_pass ? true : first_pdl, d.dp_pass ? 1u : 0u);
}

}  // namespace pgf

using namespace pgf;

extern "C" {

int pgf_step_state_set(void* state, long long noise_offset, long long t_dp, long long t_model, long long cursor, float lr,
                       float beta1, float beta2, void* stream) {
  PGF_CHECK_ARG(state && (reinterpret_cast<uintptr_t>(state) & 7) == 0, "pgf_step_state_set: state must be an 8-byte aligned device pointer");
  PGF_CHECK_ARG(t_dp >= 0 && t_model >= 0 && cursor >= 0, "pgf_step_state_set: negative step count / cursor");
  step_state_set_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<StepState*>(state), noise_offset, t_dp, t_model,
                                                                       cursor, lr, beta1, beta2);
  PGF_CUDA_LAUNCH_CHECK("pgf_step_state_set");
  return PGF_OK;
}

size_t pgf_sweep_plan_workspace(int n_models, int B, int D, int H) {
  if (n_models <= 0 || B <= 0 || D <= 0 || H <= 0) return 0;
  return SweepPlan::workspace_bytes(n_models, B, D, H, nullptr);
}

int pgf_sweep_plan_create(const pgf_sweep_desc* desc, void** plan_out) {
  PGF_CHECK_ARG(desc && plan_out, "pgf_sweep_plan_create: NULL argument");
  const pgf_sweep_desc& d = *desc;
  PGF_CHECK_ARG(d.n_models > 0 && d.B > 0 && d.B <= 8, "pgf_sweep_plan_create: the fused step is the reference-batch path, 1 <= B <= 8 (got %d)", d.B);
  PGF_CHECK_ARG(d.d0 > 0 && d.d1 >= 0 && d.d2 >= 0 && d.H > 0 && (d.H % 4) == 0 && (d.d0 % 4) == 0 && (d.d1 % 4) == 0 && (d.d2 % 4) == 0,
                "pgf_sweep_plan_create: block widths and H must be positive multiples of 4");
  PGF_CHECK_ARG(d.x0 && (d.d1 == 0 || d.x1) && (d.d2 == 0 || d.x2) && d.labels, "pgf_sweep_plan_create: dataset pointers missing");
  PGF_CHECK_ARG(d.params && d.adam_m && d.adam_v && d.grads && d.DP && d.DP_m && d.DP_v && d.dDP && d.coef && d.exp_eps && d.seeds,
                "pgf_sweep_plan_create: parameter / optimiser pointers missing");
  PGF_CHECK_ARG(d.stats_dp && d.stats_model && d.state && d.workspace, "pgf_sweep_plan_create: stats / state / workspace missing");
  PGF_CHECK_ARG((d.P % 4) == 0 && (d.off_W1 % 4) == 0 && (d.off_W2 % 4) == 0 && (d.off_Wc % 4) == 0,
                "pgf_sweep_plan_create: parameter segments must be 16-byte aligned");
  PGF_CHECK_ARG(d.tau > 0.f, "pgf_sweep_plan_create: tau must be > 0");
  const int D = d.d0 + d.d1 + d.d2;
  size_t offs[10];
  const size_t need = SweepPlan::workspace_bytes(d.n_models, d.B, D, d.H, offs);
  if (d.workspace_bytes < need) {
    set_error("pgf_sweep_plan_create: workspace of %zu bytes, need %zu", d.workspace_bytes, need);
    return PGF_ERR_WORKSPACE;
  }
  PGF_CHECK_ARG((reinterpret_cast<uintptr_t>(d.workspace) & 255) == 0, "pgf_sweep_plan_create: workspace must be 256-byte aligned");
  SweepPlan* p = new (std::nothrow) SweepPlan();
  PGF_CHECK_ARG(p, "pgf_sweep_plan_create: out of host memory");
  p->d = d;
  p->D = D;
  char* ws = static_cast<char*>(d.workspace);
  p->X = reinterpret_cast<float*>(ws + offs[0]);
  p->H1 = reinterpret_cast<float*>(ws + offs[1]);
  p->H2 = reinterpret_cast<float*>(ws + offs[2]);
  p->dZ2 = reinterpret_cast<float*>(ws + offs[3]);
  p->dZ1 = reinterpret_cast<float*>(ws + offs[4]);
  p->dX = reinterpret_cast<float*>(ws + offs[5]);
  p->dx_part = reinterpret_cast<float*>(ws + offs[6]);
  p->ce_ws = reinterpret_cast<float*>(ws + offs[7]);
  p->counters = reinterpret_cast<unsigned int*>(ws + offs[8]);
  p->dx_part_bytes = offs[9];
  p->ce_ws_bytes = cls_ce_workspace(d.B, d.H, d.n_models);
  p->pdl = d.use_pdl != 0;
  p->exec = nullptr;
  p->graph_steps = 0;
  p->graph_stream = nullptr;
  *plan_out = p;
  return PGF_OK;
}

// zero the dX counters once (they return to zero after every launch); call on the stream the plan will run on
int pgf_sweep_plan_reset(void* plan, void* stream) {
  PGF_CHECK_ARG(plan, "pgf_sweep_plan_reset: NULL plan");
  SweepPlan* p = static_cast<SweepPlan*>(plan);
  const size_t n = (static_cast<size_t>(linear_dx_counters(p->d.B, p->D, p->d.n_models)) + 1) * sizeof(unsigned int);
  PGF_CUDA_CALL(cudaMemsetAsync(p->counters, 0, n, static_cast<cudaStream_t>(stream)));
  return PGF_OK;
}

int pgf_sweep_plan_capture(void* plan, void* stream, int steps_per_graph) {
  PGF_CHECK_ARG(plan && steps_per_graph >= 1 && steps_per_graph <= 64, "pgf_sweep_plan_capture: bad argument");
  SweepPlan* p = static_cast<SweepPlan*>(plan);
  (void)stream;
  if (p->exec) {
    cudaGraphExecDestroy(p->exec);
    p->exec = nullptr;
  }
  // capture on a private stream: the legacy default stream (what a torch program runs on unless told otherwise) cannot be
  // captured, and the instantiated graph is not tied to the stream it was recorded on
  cudaStream_t cs = nullptr;
  PGF_CUDA_CALL(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  cudaError_t eb = cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed);
  if (eb != cudaSuccess) {
    cudaStreamDestroy(cs);
    set_error("pgf_sweep_plan_capture: cudaStreamBeginCapture failed: %s", cudaGetErrorString(eb));
    return PGF_ERR_CUDA;
  }
  int rc = PGF_OK;
  for (int i = 0; i < steps_per_graph && rc == PGF_OK; ++i) rc = p->enqueue(cs, i > 0);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(cs, &graph);
  cudaStreamDestroy(cs);
  if (rc != PGF_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess || !graph) {
    set_error("pgf_sweep_plan_capture: cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    return PGF_ERR_CUDA;
  }
  const cudaError_t ei = cudaGraphInstantiate(&p->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess) {
    p->exec = nullptr;
    set_error("pgf_sweep_plan_capture: cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
    return PGF_ERR_CUDA;
  }
  p->graph_steps = steps_per_graph;
  p->graph_stream = nullptr;
  return PGF_OK;
}

int pgf_sweep_plan_run(void* plan, void* stream, int n_steps) {
  PGF_CHECK_ARG(plan && n_steps >= 0, "pgf_sweep_plan_run: bad argument");
  SweepPlan* p = static_cast<SweepPlan*>(plan);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int left = n_steps;
  while (p->exec && left >= p->graph_steps) {
    PGF_CUDA_CALL(cudaGraphLaunch(p->exec, s));
    left -= p->graph_steps;
  }
  for (; left > 0; --left) {
    const int rc = p->enqueue(s, false);
    if (rc != PGF_OK) return rc;
  }
  return PGF_OK;
}

int pgf_sweep_plan_launches_per_step(void* plan) {
  if (!plan) return 0;
  const SweepPlan* p = static_cast<const SweepPlan*>(plan);
  if (p->d.n_models >= PGF_WIDE_MODELS)   // slab kernels: dX = kernel + finalize, gradient+Adam = one launch per layer + the state advance
    return p->d.dp_pass ? 18 : 9;
  return p->d.dp_pass ? 13 : 6;
}

int pgf_sweep_plan_destroy(void* plan) {
  if (!plan) return PGF_OK;
  SweepPlan* p = static_cast<SweepPlan*>(plan);
  if (p->exec) cudaGraphExecDestroy(p->exec);
  delete p;
  return PGF_OK;
}

}  // extern "C"
