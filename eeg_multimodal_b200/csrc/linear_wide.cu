// Many-models-per-GPU variants of the weight-streaming kernels of the batch-8 sweep (kernel (b), fp32 CUDA-core path).
//
// Reference ops replaced: autograd of nn.Linear in `fc_layers` (dX = dY W) and that layer's weight gradient + Adam step
// (models.py:46-51,80; past_acc.py:202-212), for every model of a sweep in one launch -- the same contracts as the
// TMA-fed persistent kernels of linear_stream.cu (pgf_linear_bwd_dx / pgf_linear_adam_step dispatch between the two).
//
// linear_stream.cu is built for FEW models per GPU (6 of the 48-model sweep on each of 8 GPUs): one CTA per SM walks
// 512-byte column chunks through a TMA ring, which hides the ramp-up and tail of 20-150 us kernels.  With the 48 or 128
// models per GPU of the single-GPU sweep and of the 1024-model ablation grid (BASELINE config 5) a kernel runs for
// milliseconds, ramp-up and tail do not matter, and what decides the bandwidth is how many 2 MB pages the CTAs resident
// at one moment stream from: here one CTA owns a SHORT slab of full-width weight rows (32 rows for gradient+Adam, <= 128
// for dX) and the grid is launched model-major, so the resident CTAs sweep a few contiguous tens of MB of W / m / v.
// Measured with 128 models per GPU (profiles/README.md): gradient+Adam of the 2304x2304 layer 3.01 ms here against
// 3.26 ms for the ring kernel, dX 0.50 against 0.64 ms.  Same per-element arithmetic (adam_update, fixed-order FMAs).
#include <stdlib.h>

#include "pgf_kernels.cuh"

namespace pgf {

#define PGF_ACT_NONE 0
#define PGF_ACT_RELU 1
#define PGF_ACT_TANH 2

constexpr int TB = 8;  // batch rows per tile (the reference's batch size)

// ------------------------------------------------------------------------------------------
// dX
// ------------------------------------------------------------------------------------------
struct WideDxArgs {
  const float* dY; long long ldy; long long sdY;  // [B,N]
  const float* W; long long sW;                   // [N,K]
  float* partial;                                 // [n_models][bchunks][nslab][TB][K]
  int B, N, K, nslab, rows_per_slab;
};

__global__ void __launch_bounds__(128) wide_dx_kernel(const WideDxArgs a) {
  extern __shared__ float sdy[];  // [rows_per_slab][TB]
  const int bchunks = (a.B + TB - 1) / TB;
  const int model = blockIdx.z / bchunks, bc = blockIdx.z - model * bchunks;
  const int b0 = bc * TB, nb = min(TB, a.B - b0);
  const int slab = blockIdx.y;
  const int n0 = slab * a.rows_per_slab, n1 = min(a.N, n0 + a.rows_per_slab);
  const float* dY = a.dY + model * a.sdY + static_cast<long long>(b0) * a.ldy;
  for (int i = threadIdx.x; i < (n1 - n0) * TB; i += blockDim.x) {
    const int n = i / TB, b = i - n * TB;
    sdy[i] = b < nb ? dY[b * a.ldy + n0 + n] : 0.f;
  }
  __syncthreads();
  const int k4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (k4 * 4 >= a.K) return;
  const float4* W = reinterpret_cast<const float4*>(a.W + model * a.sW) + k4;
  const int K4 = a.K >> 2;
  float4 acc[TB];
#pragma unroll
  for (int b = 0; b < TB; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int U = 8;  // weight rows in flight per thread
  int n = n0;
  for (; n + U <= n1; n += U) {
    float4 w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) w[u] = ldg_stream(W + static_cast<long long>(n + u) * K4);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 g0 = *reinterpret_cast<const float4*>(sdy + (n - n0 + u) * TB);
      const float4 g1 = *reinterpret_cast<const float4*>(sdy + (n - n0 + u) * TB + 4);
      const float g[TB] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int b = 0; b < TB; ++b) {
        acc[b].x = fmaf(g[b], w[u].x, acc[b].x);
        acc[b].y = fmaf(g[b], w[u].y, acc[b].y);
        acc[b].z = fmaf(g[b], w[u].z, acc[b].z);
        acc[b].w = fmaf(g[b], w[u].w, acc[b].w);
      }
    }
  }
  for (; n < n1; ++n) {
    const float4 w = ldg_stream(W + static_cast<long long>(n) * K4);
    const float4 g0 = *reinterpret_cast<const float4*>(sdy + (n - n0) * TB);
    const float4 g1 = *reinterpret_cast<const float4*>(sdy + (n - n0) * TB + 4);
    const float g[TB] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int b = 0; b < TB; ++b) {
      acc[b].x = fmaf(g[b], w.x, acc[b].x);
      acc[b].y = fmaf(g[b], w.y, acc[b].y);
      acc[b].z = fmaf(g[b], w.z, acc[b].z);
      acc[b].w = fmaf(g[b], w.w, acc[b].w);
    }
  }
  float4* P = reinterpret_cast<float4*>(a.partial) +
              ((static_cast<long long>(blockIdx.z) * a.nslab + slab) * TB) * K4 + k4;
#pragma unroll
  for (int b = 0; b < TB; ++b) P[static_cast<long long>(b) * K4] = acc[b];
}

// sum the slab partials; optionally apply the derivative of the activation that produced `mask_src`
// (the layer input): RELU -> * (src > 0), TANH -> * (1 - src^2)
__global__ void wide_dx_finalize_kernel(const float* __restrict__ partial, int nslab, int B, int K, int bchunks,
                                          const float* __restrict__ mask_src, int mask_mode, long long ld_mask,
                                          long long s_mask, float* __restrict__ dX, long long ldx, long long sdX) {
  const int K4 = K >> 2;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int model = blockIdx.y;
  if (idx >= static_cast<long long>(B) * K4) return;
  const int b = static_cast<int>(idx / K4), k4 = static_cast<int>(idx - static_cast<long long>(b) * K4);
  const int bc = b / TB, bl = b - bc * TB;
  const float4* P = reinterpret_cast<const float4*>(partial) +
                    (((static_cast<long long>(model) * bchunks + bc) * nslab) * TB + bl) * K4 + k4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < nslab; ++i) {
    const float4 p = P[static_cast<long long>(i) * TB * K4];
    s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
  }
  if (mask_src) {
    const float4 m = *reinterpret_cast<const float4*>(mask_src + model * s_mask + static_cast<long long>(b) * ld_mask + 4 * k4);
    if (mask_mode == PGF_ACT_TANH) {
      s.x *= 1.f - m.x * m.x;
      s.y *= 1.f - m.y * m.y;
      s.z *= 1.f - m.z * m.z;
      s.w *= 1.f - m.w * m.w;
    } else {
      s.x = m.x > 0.f ? s.x : 0.f;
      s.y = m.y > 0.f ? s.y : 0.f;
      s.z = m.z > 0.f ? s.z : 0.f;
      s.w = m.w > 0.f ? s.w : 0.f;
    }
  }
  *reinterpret_cast<float4*>(dX + model * sdX + static_cast<long long>(b) * ldx + 4 * k4) = s;
}

static int wide_dx_slabs(int B, int N, int K, int n_models) {
  const int kctas = (K / 4 + 127) / 128;
  const int bchunks = (B + TB - 1) / TB;
  const long long base = static_cast<long long>(kctas) * bchunks * n_models;
  int slabs = static_cast<int>((8LL * num_sms() + base - 1) / base);  // ~8 CTAs of 4 warps per SM
  // ... but never slabs longer than ~128 weight rows: short slabs in model-major launch order keep the set of 2 MB
  // pages the resident CTAs touch small (measured 0.67 -> 0.77 of HBM peak on the 2304x2304 layer)
  static const int rows_env = getenv("PGF_LINDX_ROWS") ? atoi(getenv("PGF_LINDX_ROWS")) : 0;
  const int rows_target = rows_env > 0 ? rows_env : (N >= 2048 ? 128 : 64);
  const int by_rows = (N + rows_target - 1) / rows_target;
  if (by_rows > slabs) slabs = by_rows;
  if (slabs < 1) slabs = 1;
  const int max_slabs = (N + 15) / 16;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs > 64) slabs = 64;
  return slabs;
}

size_t linear_dx_workspace_wide(int B, int N, int K, int n_models) {
  const int bchunks = (B + TB - 1) / TB;
  return static_cast<size_t>(n_models) * bchunks * wide_dx_slabs(B, N, K, n_models) * TB * K * sizeof(float);
}

int linear_bwd_dx_wide(const float* dY, long long ldy, long long sdY, const float* W, long long sW, const float* mask_src,
                  int mask_mode, long long ld_mask, long long s_mask, float* dX, long long ldx, long long sdX, int B, int N, int K,
                  int n_models, float* workspace, size_t workspace_bytes, cudaStream_t s) {
  if (workspace_bytes < linear_dx_workspace_wide(B, N, K, n_models)) {
    set_error("pgf_linear_bwd_dx: workspace too small");
    return PGF_ERR_WORKSPACE;
  }
  WideDxArgs a;
  a.dY = dY; a.ldy = ldy; a.sdY = sdY; a.W = W; a.sW = sW; a.partial = workspace;
  a.B = B; a.N = N; a.K = K;
  a.nslab = wide_dx_slabs(B, N, K, n_models);
  a.rows_per_slab = (N + a.nslab - 1) / a.nslab;
  const int bchunks = (B + TB - 1) / TB;
  const dim3 grid((K / 4 + 127) / 128, a.nslab, n_models * bchunks);
  const size_t smem = static_cast<size_t>(a.rows_per_slab) * TB * sizeof(float);
  ensure_dynamic_smem(reinterpret_cast<const void*>(wide_dx_kernel), smem);
  wide_dx_kernel<<<grid, 128, smem, s>>>(a);
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_bwd_dx");
  const long long total = static_cast<long long>(B) * (K / 4);
  const dim3 fgrid(static_cast<unsigned>((total + 255) / 256), n_models);
  wide_dx_finalize_kernel<<<fgrid, 256, 0, s>>>(workspace, a.nslab, B, K, bchunks, mask_src, mask_mode, ld_mask, s_mask, dX, ldx, sdX);
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_bwd_dx(finalize)");
  return PGF_OK;
}


// ------------------------------------------------------------------------------------------
// dW fused into Adam: dW[n,k] = sum_b dY[b,n] X[b,k] is a rank-B outer product, recomputed per element inside the optimiser
// (8 FMAs) instead of being written to HBM by one kernel and read back by the next: 24 instead of 32 bytes per parameter.
// (Keeping the activation tile in shared memory instead of 32 registers per thread was tried for occupancy: the
// extra LDS traffic and the spills at 72 registers made it 1.8x slower; 5 CTAs/SM with x in registers it is.)
// ------------------------------------------------------------------------------------------
__global__ void step_advance_kernel(const StepAdvance v) { step_advance_apply(v); }

struct WideAdamArgs {
  const float* dY; long long ldy; long long sdY;   // [B,N] output gradient
  const float* X; long long ldx; long long sX;     // [B,K] layer input
  float* W; float* mW; float* vW;                  // [N,K] weight and its Adam moments
  float* bias; float* mb; float* vb;               // [N] (optional)
  long long sP;                                    // model stride of W/mW/vW/bias/mb/vb (one flat buffer per model)
  int B, N, K, rows_per_cta;
  AdamCoef c;
  const StepState* st;                             // step-dependent Adam coefficients from the device state (optional)
};

template <int U>  // U rows in flight per thread (3 x 128-bit loads each)
__global__ void __launch_bounds__(128, 5) wide_adam_kernel(const WideAdamArgs a) {
  extern __shared__ float sdy[];  // [rows_per_cta][TB]
  const AdamCoef coef = adam_coef_at(a.c, a.st, 1);
  const int model = blockIdx.z;
  const int n0 = blockIdx.y * a.rows_per_cta, n1 = min(a.N, n0 + a.rows_per_cta);
  const int k4 = blockIdx.x * blockDim.x + threadIdx.x;
  const int K4 = a.K >> 2;
  const float* dY = a.dY + model * a.sdY;
  for (int i = threadIdx.x; i < (n1 - n0) * TB; i += blockDim.x) {
    const int n = i / TB, b = i - n * TB;
    sdy[i] = b < a.B ? dY[b * a.ldy + n0 + n] : 0.f;
  }
  __syncthreads();
  if (a.bias && blockIdx.x == 0) {  // bias: gradient = column sum of dY
    for (int n = threadIdx.x; n < n1 - n0; n += blockDim.x) {
      float g = 0.f;
#pragma unroll
      for (int b = 0; b < TB; ++b) g += sdy[n * TB + b];
      const long long i = model * a.sP + n0 + n;
      float p = a.bias[i], m = a.mb[i], v = a.vb[i];
      adam_update(p, m, v, g, coef);
      a.bias[i] = p; a.mb[i] = m; a.vb[i] = v;
    }
  }
  if (k4 >= K4) return;
  float4 x[TB];
  const float* X = a.X + model * a.sX + 4 * k4;
#pragma unroll
  for (int b = 0; b < TB; ++b) x[b] = b < a.B ? *reinterpret_cast<const float4*>(X + b * a.ldx) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4* W = reinterpret_cast<float4*>(a.W + model * a.sP) + k4;
  float4* M = reinterpret_cast<float4*>(a.mW + model * a.sP) + k4;
  float4* V = reinterpret_cast<float4*>(a.vW + model * a.sP) + k4;
  // one row: gradient from the rank-8 outer product, Adam update in registers
  auto update_row = [&](int n, float4& p, float4& m, float4& v) {
    const float4 g0 = *reinterpret_cast<const float4*>(sdy + (n - n0) * TB);
    const float4 g1 = *reinterpret_cast<const float4*>(sdy + (n - n0) * TB + 4);
    const float g[TB] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int b = 0; b < TB; ++b) {   // same order as linear_dw_kernel: bit-identical gradient
      o.x = fmaf(g[b], x[b].x, o.x);
      o.y = fmaf(g[b], x[b].y, o.y);
      o.z = fmaf(g[b], x[b].z, o.z);
      o.w = fmaf(g[b], x[b].w, o.w);
    }
    adam_update(p.x, m.x, v.x, o.x, coef);
    adam_update(p.y, m.y, v.y, o.y, coef);
    adam_update(p.z, m.z, v.z, o.z, coef);
    adam_update(p.w, m.w, v.w, o.w, coef);
    const long long off = static_cast<long long>(n) * K4;
    W[off] = p; M[off] = m; V[off] = v;
  };
  // Two register sets in ping-pong, written out by hand (no copies between them): set B's loads are issued before
  // set A is consumed and vice versa, so a set's scoreboard wait never covers the loads issued after it.
  float4 pA[U], mA[U], vA[U], pB[U], mB[U], vB[U];
  auto load_set = [&](int n, float4 (&p)[U], float4 (&m)[U], float4 (&v)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (n + u < n1) {
        const long long o = static_cast<long long>(n + u) * K4;
        p[u] = W[o]; m[u] = M[o]; v[u] = V[o];
      }
    }
  };
  auto do_set = [&](int n, float4 (&p)[U], float4 (&m)[U], float4 (&v)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (n + u < n1) update_row(n + u, p[u], m[u], v[u]);
  };
  int n = n0;
  load_set(n, pA, mA, vA);
  for (; n < n1; n += 2 * U) {
    load_set(n + U, pB, mB, vB);
    do_set(n, pA, mA, vA);
    load_set(n + 2 * U, pA, mA, vA);
    do_set(n + U, pB, mB, vB);
  }
}

int linear_adam_step_wide(const LinAdamArgs& in, int n_models, cudaStream_t s) {
  if (in.B > TB) {
    set_error("pgf_linear_adam_step: the fused gradient+Adam kernel handles batches up to %d rows (got %d)", TB, in.B);
    return PGF_ERR_UNSUPPORTED;
  }
  // Small row slabs, model-major launch order: the CTAs resident at any moment then cover a few contiguous tens of MB
  // of W / m / v instead of one long stream per CTA scattered over the whole 4 GB sweep state -- with 467-row slabs
  // the kernel ran at 73 % of HBM peak, with 32-row slabs at 85 % (fewer 2 MB pages live at once).
  static const int rows_exact = getenv("PGF_LINADAM_ROWS") ? atoi(getenv("PGF_LINADAM_ROWS")) : 0;
  for (int i = 0; i < in.n_layers; ++i) {
    const LinAdamLayer& l = in.l[i];
    WideAdamArgs a;
    a.dY = l.dY; a.ldy = l.ldy; a.sdY = l.sdY; a.X = l.X; a.ldx = l.ldx; a.sX = l.sX; a.W = l.W; a.mW = l.mW; a.vW = l.vW;
    a.bias = l.bias; a.mb = l.mb; a.vb = l.vb; a.sP = in.sP; a.B = in.B; a.N = l.N; a.K = l.K; a.c = in.c; a.st = in.st;
    int rows = rows_exact > 0 ? rows_exact : 32;
    if (rows > a.N) rows = a.N;
    a.rows_per_cta = rows;
    const int kctas = (a.K / 4 + 127) / 128;
    const dim3 grid(kctas, (a.N + rows - 1) / rows, n_models);
    const size_t smem = static_cast<size_t>(rows) * TB * sizeof(float);
    wide_adam_kernel<1><<<grid, 128, smem, s>>>(a);  // two rows per register set spill at the 5-CTAs/SM register budget
    PGF_CUDA_LAUNCH_CHECK("pgf_linear_adam_step");
  }
  if (in.adv.st) {   // sweep step: every kernel above has read the device step state; a last one-thread launch advances it
    step_advance_kernel<<<1, 1, 0, s>>>(in.adv);
    PGF_CUDA_LAUNCH_CHECK("pgf_linear_adam_step(step advance)");
  }
  return PGF_OK;
}

}  // namespace pgf
