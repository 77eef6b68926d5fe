// fp32 CUDA-core path of the fusion MLP (kernel (b), small-batch / parity regime).
//
// Reference ops replaced: nn.Linear + ReLU / Tanh of `fc_layers` and their autograd
// (models.py:46-51,80; past_acc.py:87-92,137).  At the reference's batch size (B=8) a layer is
// [8,K]x[K,N]: ~8 flop per weight byte, i.e. bound by streaming the fp32 weights from HBM, not
// by math.  So these kernels are organised around ONE coalesced pass over W per call, with the
// 8 activation rows held in shared memory / registers, and they are GROUPED: blockIdx.z walks
// the independent models of an eps x seed sweep (strided batched layout), so that a sweep fills
// the 148 SMs even though one model cannot.
//   fwd : Y[b,n]  = act(sum_k X[b,k] W[n,k] + bias[n])         read W once
//   dx  : dX[b,k] = sum_n dY[b,n] W[n,k]  (* relu mask)        read W once
//   dw  : dW[n,k] = sum_b dY[b,n] X[b,k],  db[n] = sum_b dY    write dW once
// Accumulation is fp32 FMA in a fixed order (deterministic; no atomics).
#include <stdlib.h>

#include "pgf_kernels.cuh"

namespace pgf {

#define PGF_ACT_NONE 0
#define PGF_ACT_RELU 1
#define PGF_ACT_TANH 2

constexpr int TB = 8;  // batch rows per tile (the reference's batch size)


template <int R>
__global__ void __launch_bounds__(256) linear_fwd_kernel(const LinFwdArgs a) {
  extern __shared__ float4 sx4[];  // [TB][K/4]
  const int model = blockIdx.z;
  const int b0 = blockIdx.y * TB;
  const int nb = min(TB, a.B - b0);
  const int K4 = a.K >> 2;
  const float* X = a.X + model * a.sX + static_cast<long long>(b0) * a.ldx;
  const float* W = a.W + model * a.sW;
  // Activation tile -> shared memory with 16-byte async copies (LDGSTS: no register staging, all of a thread's copies
  // in flight at once); the first weight vectors are requested BEFORE the wait, so the weight stream is already running
  // while the tile lands (ncu: the LDG->STS staging loop held 23 % of this kernel's stall samples).
  for (int i = threadIdx.x; i < TB * K4; i += blockDim.x) {
    const int b = i / K4, k = i - b * K4;
    if (b < nb) {
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(sx4 + i));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(X + b * a.ldx + 4 * k) : "memory");
    } else {
      sx4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n0 = (blockIdx.x * (blockDim.x >> 5) + warp) * R;
  const bool active = n0 < a.N;
  float acc[R * TB];
#pragma unroll
  for (int i = 0; i < R * TB; ++i) acc[i] = 0.f;
  const float4* wrow[R];
#pragma unroll
  for (int r = 0; r < R; ++r) wrow[r] = reinterpret_cast<const float4*>(W + static_cast<long long>(min(n0 + r, a.N - 1)) * a.K);

  // Two register sets of R weight vectors in ping-pong (written out by hand, no copies between them): the loads of
  // k-step i+1 are issued before k-step i is consumed, and a set's scoreboard wait never covers the other set's loads.
  auto load_w = [&](int k, float4 (&w)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (k < K4) w[r] = ldg_stream(wrow[r] + k);
  };
  auto consume = [&](int k, const float4 (&w)[R]) {
    if (k >= K4) return;
#pragma unroll
    for (int b = 0; b < TB; ++b) {
      const float4 x = sx4[b * K4 + k];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float s = acc[r * TB + b];
        s = fmaf(x.x, w[r].x, s);
        s = fmaf(x.y, w[r].y, s);
        s = fmaf(x.z, w[r].z, s);
        s = fmaf(x.w, w[r].w, s);
        acc[r * TB + b] = s;
      }
    }
  };
  float4 wA[R], wB[R];
  int k = lane;
  if (active) load_w(k, wA);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (!active) return;
  for (; k < K4; k += 64) {
    load_w(k + 32, wB);
    consume(k, wA);
    load_w(k + 64, wA);
    consume(k + 32, wB);
  }
  // warp reduction of R*TB values; every lane ends with the full sums
#pragma unroll
  for (int i = 0; i < R * TB; ++i) acc[i] = warp_sum(acc[i]);
  // lane i < R*TB writes value i  (i = r*TB + b)
  float mine = 0.f;
#pragma unroll
  for (int i = 0; i < R * TB; ++i)
    if (lane == i) mine = acc[i];
  if (lane < R * TB) {
    const int r = lane / TB, b = lane - r * TB;
    const int n = n0 + r;
    if (n < a.N && b < nb) {
      float v = mine + (a.bias ? a.bias[model * a.sb + n] : 0.f);
      if (a.act == PGF_ACT_RELU) v = fmaxf(v, 0.f);
      else if (a.act == PGF_ACT_TANH) v = tanhf(v);
      a.Y[model * a.sY + static_cast<long long>(b0 + b) * a.ldy + n] = v;
    }
  }
}

int linear_fwd(const LinFwdArgs& a, int n_models, cudaStream_t s) {
  const size_t smem = static_cast<size_t>(TB) * a.K * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("pgf_linear_fwd: K=%d too large for the shared-memory activation tile", a.K);
    return PGF_ERR_UNSUPPORTED;
  }
  const int warps = 8;
  const int bchunks = (a.B + TB - 1) / TB;
  const long long ctas_r4 = static_cast<long long>((a.N + warps * 4 - 1) / (warps * 4)) * bchunks * n_models;
  const dim3 block(warps * 32);
  if (ctas_r4 >= 2LL * num_sms()) {
    ensure_dynamic_smem(reinterpret_cast<const void*>(linear_fwd_kernel<4>), smem);
    const dim3 grid((a.N + warps * 4 - 1) / (warps * 4), bchunks, n_models);
    linear_fwd_kernel<4><<<grid, block, smem, s>>>(a);
  } else {
    ensure_dynamic_smem(reinterpret_cast<const void*>(linear_fwd_kernel<2>), smem);
    const dim3 grid((a.N + warps * 2 - 1) / (warps * 2), bchunks, n_models);
    linear_fwd_kernel<2><<<grid, block, smem, s>>>(a);
  }
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_fwd");
  return PGF_OK;
}

// ------------------------------------------------------------------------------------------
// dX
// ------------------------------------------------------------------------------------------
struct LinDxArgs {
  const float* dY; long long ldy; long long sdY;  // [B,N]
  const float* W; long long sW;                   // [N,K]
  float* partial;                                 // [n_models][bchunks][nslab][TB][K]
  int B, N, K, nslab, rows_per_slab;
};

__global__ void __launch_bounds__(128) linear_dx_kernel(const LinDxArgs a) {
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
__global__ void linear_dx_finalize_kernel(const float* __restrict__ partial, int nslab, int B, int K, int bchunks,
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

int linear_dx_slabs(int B, int N, int K, int n_models) {
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

size_t linear_dx_workspace(int B, int N, int K, int n_models) {
  const int bchunks = (B + TB - 1) / TB;
  return static_cast<size_t>(n_models) * bchunks * linear_dx_slabs(B, N, K, n_models) * TB * K * sizeof(float);
}

int linear_bwd_dx(const float* dY, long long ldy, long long sdY, const float* W, long long sW, const float* mask_src,
                  int mask_mode, long long ld_mask, long long s_mask, float* dX, long long ldx, long long sdX, int B, int N, int K,
                  int n_models, float* workspace, size_t workspace_bytes, cudaStream_t s) {
  if (workspace_bytes < linear_dx_workspace(B, N, K, n_models)) {
    set_error("pgf_linear_bwd_dx: workspace too small");
    return PGF_ERR_WORKSPACE;
  }
  LinDxArgs a;
  a.dY = dY; a.ldy = ldy; a.sdY = sdY; a.W = W; a.sW = sW; a.partial = workspace;
  a.B = B; a.N = N; a.K = K;
  a.nslab = linear_dx_slabs(B, N, K, n_models);
  a.rows_per_slab = (N + a.nslab - 1) / a.nslab;
  const int bchunks = (B + TB - 1) / TB;
  const dim3 grid((K / 4 + 127) / 128, a.nslab, n_models * bchunks);
  const size_t smem = static_cast<size_t>(a.rows_per_slab) * TB * sizeof(float);
  ensure_dynamic_smem(reinterpret_cast<const void*>(linear_dx_kernel), smem);
  linear_dx_kernel<<<grid, 128, smem, s>>>(a);
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_bwd_dx");
  const long long total = static_cast<long long>(B) * (K / 4);
  const dim3 fgrid(static_cast<unsigned>((total + 255) / 256), n_models);
  linear_dx_finalize_kernel<<<fgrid, 256, 0, s>>>(workspace, a.nslab, B, K, bchunks, mask_src, mask_mode, ld_mask, s_mask, dX, ldx, sdX);
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_bwd_dx(finalize)");
  return PGF_OK;
}

// ------------------------------------------------------------------------------------------
// dW, db
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) linear_dw_kernel(const LinDwArgs a) {
  extern __shared__ float sdy[];  // [rows_per_cta][TB]
  const int model = blockIdx.z;
  const int n0 = blockIdx.y * a.rows_per_cta, n1 = min(a.N, n0 + a.rows_per_cta);
  const int k4 = blockIdx.x * blockDim.x + threadIdx.x;
  const int K4 = a.K >> 2;
  const bool active = k4 < K4;
  float4* dW = reinterpret_cast<float4*>(a.dW + model * a.sdW) + k4;
  for (int b0 = 0; b0 < a.B; b0 += TB) {
    const int nb = min(TB, a.B - b0);
    const float* dY = a.dY + model * a.sdY + static_cast<long long>(b0) * a.ldy;
    __syncthreads();
    for (int i = threadIdx.x; i < (n1 - n0) * TB; i += blockDim.x) {
      const int n = i / TB, b = i - n * TB;
      sdy[i] = b < nb ? dY[b * a.ldy + n0 + n] : 0.f;
    }
    __syncthreads();
    if (a.db && blockIdx.x == 0) {  // bias gradient: one thread per row
      for (int n = threadIdx.x; n < n1 - n0; n += blockDim.x) {
        float sum = 0.f;
#pragma unroll
        for (int b = 0; b < TB; ++b) sum += sdy[n * TB + b];
        float* p = a.db + model * a.sdb + n0 + n;
        *p = (b0 > 0 || a.accumulate) ? *p + sum : sum;
      }
    }
    if (!active) continue;
    float4 x[TB];
    const float* X = a.X + model * a.sX + static_cast<long long>(b0) * a.ldx + 4 * k4;
#pragma unroll
    for (int b = 0; b < TB; ++b) x[b] = b < nb ? *reinterpret_cast<const float4*>(X + b * a.ldx) : make_float4(0.f, 0.f, 0.f, 0.f);
    const bool acc_mode = (b0 > 0) || a.accumulate;
    for (int n = n0; n < n1; ++n) {
      const float4 g0 = *reinterpret_cast<const float4*>(sdy + (n - n0) * TB);
      const float4 g1 = *reinterpret_cast<const float4*>(sdy + (n - n0) * TB + 4);
      const float g[TB] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      float4 o = acc_mode ? dW[static_cast<long long>(n) * K4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int b = 0; b < TB; ++b) {
        o.x = fmaf(g[b], x[b].x, o.x);
        o.y = fmaf(g[b], x[b].y, o.y);
        o.z = fmaf(g[b], x[b].z, o.z);
        o.w = fmaf(g[b], x[b].w, o.w);
      }
      dW[static_cast<long long>(n) * K4] = o;
    }
  }
}

int linear_bwd_dw(const LinDwArgs& a_in, int n_models, cudaStream_t s) {
  LinDwArgs a = a_in;
  const int kctas = (a.K / 4 + 127) / 128;
  // enough row-blocks to fill the GPU, but >= 16 rows each so the dY tile load is amortised
  long long want = (4LL * num_sms() + static_cast<long long>(kctas) * n_models - 1) / (static_cast<long long>(kctas) * n_models);
  if (want < 1) want = 1;
  int rows = static_cast<int>((a.N + want - 1) / want);
  if (rows < 16) rows = 16;
  if (rows > 1024) rows = 1024;
  a.rows_per_cta = rows;
  const dim3 grid(kctas, (a.N + rows - 1) / rows, n_models);
  const size_t smem = static_cast<size_t>(rows) * TB * sizeof(float);
  linear_dw_kernel<<<grid, 128, smem, s>>>(a);
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_bwd_dw");
  return PGF_OK;
}

// ------------------------------------------------------------------------------------------
// dW fused into Adam (small batch): dW[n,k] = sum_b dY[b,n] X[b,k] is a rank-B outer product, so at the
// reference batch size it is recomputed inside the optimiser (8 FMAs per element) instead of being written
// to HBM by one kernel and read back by the next: 24 instead of 32 bytes per parameter and step.
// Same thread <-> element mapping as linear_dw_kernel; the arithmetic per element is adam_update().
// ------------------------------------------------------------------------------------------
// (Keeping the activation tile in shared memory instead of 32 registers per thread was tried for occupancy: the
// extra LDS traffic and the spills at 72 registers made it 1.8x slower; 5 CTAs/SM with x in registers it is.)
template <int U>  // U rows in flight per thread (3 x 128-bit loads each)
__global__ void __launch_bounds__(128, 5) linear_adam_kernel(const LinAdamArgs a) {
  extern __shared__ float sdy[];  // [rows_per_cta][TB]
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
      adam_update(p, m, v, g, a.c);
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
    adam_update(p.x, m.x, v.x, o.x, a.c);
    adam_update(p.y, m.y, v.y, o.y, a.c);
    adam_update(p.z, m.z, v.z, o.z, a.c);
    adam_update(p.w, m.w, v.w, o.w, a.c);
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

int linear_adam_step(const LinAdamArgs& a_in, int n_models, cudaStream_t s) {
  LinAdamArgs a = a_in;
  if (a.B > TB) {
    set_error("pgf_linear_adam_step: the fused gradient+Adam kernel handles batches up to %d rows (got %d)", TB, a.B);
    return PGF_ERR_UNSUPPORTED;
  }
  const int kctas = (a.K / 4 + 127) / 128;
  // Small row slabs, model-major launch order: the CTAs resident at any moment then cover a few contiguous tens of MB
  // of W / m / v instead of one long stream per CTA scattered over the whole 4 GB sweep state -- with 467-row slabs
  // the kernel ran at 73 % of HBM peak, with 32-row slabs at 85 % (fewer 2 MB pages live at once).
  static const int rows_exact = getenv("PGF_LINADAM_ROWS") ? atoi(getenv("PGF_LINADAM_ROWS")) : 0;
  int rows = rows_exact > 0 ? rows_exact : 32;
  if (rows > a.N) rows = a.N;
  a.rows_per_cta = rows;
  const dim3 grid(kctas, (a.N + rows - 1) / rows, n_models);
  const size_t smem = static_cast<size_t>(rows) * TB * sizeof(float);
  linear_adam_kernel<1><<<grid, 128, smem, s>>>(a);  // two rows per register set spill at the 5-CTAs/SM register budget
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_adam_step");
  return PGF_OK;
}

}  // namespace pgf
