// fp32 CUDA-core path of the fusion MLP (kernel (b), small-batch / parity regime).
//
// Reference ops replaced: nn.Linear + ReLU / Tanh of `fc_layers` and their autograd
// (models.py:46-51,80; past_acc.py:87-92,137).  At the reference's batch size (B=8) a layer is
// [8,K]x[K,N]: ~8 flop per weight byte, i.e. bound by streaming the fp32 weights from HBM, not
// by math.  So these kernels are organised around ONE coalesced pass over W per call, with the
// 8 activation rows held in shared memory / registers, and they are GROUPED: blockIdx.z walks
// the independent models of an eps x seed sweep (strided batched layout), so that a sweep fills
// the 148 SMs even though one model cannot.
//   dw  : dW[n,k] = sum_b dY[b,n] X[b,k],  db[n] = sum_b dY    write dW once
// (the forward, dX and fused gradient+Adam kernels of this path are the TMA-fed persistent kernels of linear_stream.cu;
//  this file keeps the materialised weight gradient, which the gradient all-reduce of the data-parallel mode needs)
// Accumulation is fp32 FMA in a fixed order (deterministic; no atomics).
#include <stdlib.h>

#include "pgf_kernels.cuh"

namespace pgf {

#define PGF_ACT_NONE 0
#define PGF_ACT_RELU 1
#define PGF_ACT_TANH 2

constexpr int TB = 8;  // batch rows per tile (the reference's batch size)


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
// dW, db of a NARROW layer at a LARGE batch (the 768 -> 2 classifier behind the module API, models.py:81, when
// ConcatModel.forward runs on a 65,536-row batch): dW[n,k] = sum_b dY[b,n] X[b,k] with N <= 8.  linear_dw_kernel walks the
// batch sequentially inside each CTA, and a 2-row layer gives it two CTAs (10 ms at B = 65,536); here the batch is cut into
// slabs, one CTA per (column block, slab) keeps its N partial rows in registers, and the slab partials are summed in slab
// order by reduce_partials (fixed order, no atomics).
// ------------------------------------------------------------------------------------------
struct LinDwBatchArgs {
  const float* dY; long long ldy, sdY;
  const float* X; long long ldx, sX;
  float* partial;      // [n_models][nslab][N][K]
  float* partial_db;   // [n_models][nslab][N]
  int B, N, K, rows_per_slab, nslab;
};

template <int NN>
__global__ void __launch_bounds__(128) linear_dw_batch_kernel(const LinDwBatchArgs a) {
  extern __shared__ float sdy[];  // [rows of the slab][N]
  const int model = blockIdx.z, slab = blockIdx.y;
  const int b0 = slab * a.rows_per_slab, b1 = min(a.B, b0 + a.rows_per_slab), rows = b1 - b0;
  const float* dY = a.dY + model * a.sdY;
  for (int i = threadIdx.x; i < rows * a.N; i += blockDim.x) {
    const int r = i / a.N, n = i - r * a.N;
    sdy[i] = dY[static_cast<long long>(b0 + r) * a.ldy + n];
  }
  __syncthreads();
  const long long unit = static_cast<long long>(model) * a.nslab + slab;
  if (blockIdx.x == 0 && a.partial_db && threadIdx.x < a.N) {
    float sum = 0.f;
    for (int r = 0; r < rows; ++r) sum += sdy[r * a.N + threadIdx.x];
    a.partial_db[unit * a.N + threadIdx.x] = sum;
  }
  const int k4 = blockIdx.x * blockDim.x + threadIdx.x, K4 = a.K >> 2;
  if (k4 >= K4) return;
  float4 acc[NN];
#pragma unroll
  for (int n = 0; n < NN; ++n) acc[n] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* X = reinterpret_cast<const float4*>(a.X + model * a.sX + static_cast<long long>(b0) * a.ldx) + k4;
  const long long ldx4 = a.ldx >> 2;
  constexpr int U = 4;  // rows in flight per thread
  int r = 0;
  for (; r + U <= rows; r += U) {
    float4 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) x[u] = ldg_stream(X + static_cast<long long>(r + u) * ldx4);
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int n = 0; n < NN; ++n) {
        if (n < a.N) {
          const float g = sdy[(r + u) * a.N + n];
          acc[n].x = fmaf(g, x[u].x, acc[n].x);
          acc[n].y = fmaf(g, x[u].y, acc[n].y);
          acc[n].z = fmaf(g, x[u].z, acc[n].z);
          acc[n].w = fmaf(g, x[u].w, acc[n].w);
        }
      }
    }
  }
  for (; r < rows; ++r) {
    const float4 x = ldg_stream(X + static_cast<long long>(r) * ldx4);
#pragma unroll
    for (int n = 0; n < NN; ++n) {
      if (n < a.N) {
        const float g = sdy[r * a.N + n];
        acc[n].x = fmaf(g, x.x, acc[n].x);
        acc[n].y = fmaf(g, x.y, acc[n].y);
        acc[n].z = fmaf(g, x.z, acc[n].z);
        acc[n].w = fmaf(g, x.w, acc[n].w);
      }
    }
  }
  float4* P = reinterpret_cast<float4*>(a.partial + unit * a.N * a.K) + k4;
#pragma unroll
  for (int n = 0; n < NN; ++n)
    if (n < a.N) P[static_cast<long long>(n) * K4] = acc[n];
}

// dX of the same narrow layer at a large batch: dX[b,k] = sum_n dY[b,n] W[n,k] (* act'(mask)) is N <= 8 FMAs per output
// element, i.e. an elementwise pass over the [B,K] output with the whole weight matrix in L1 -- not a weight-streaming
// problem (the ring kernel walks 8-row batch tiles: 0.56 ms for the classifier at B = 65,536 against 0.07 ms here).
template <int NN>
__global__ void __launch_bounds__(256) linear_dx_narrow_kernel(const float* __restrict__ dY, long long ldy, long long sdY,
                                                               const float* __restrict__ W, long long sW,
                                                               const float* __restrict__ mask_src, int mask_mode, long long ld_mask,
                                                               long long s_mask, float* __restrict__ dX, long long ldx, long long sdX,
                                                               int B, int N, int K) {
  const int model = blockIdx.y, K4 = K >> 2;
  const long long total = static_cast<long long>(B) * K4;
  const float* dYm = dY + model * sdY;
  const float4* Wm = reinterpret_cast<const float4*>(W + model * sW);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / K4), k4 = static_cast<int>(i - static_cast<long long>(b) * K4);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int n = 0; n < NN; ++n) {
      if (n < N) {
        const float g = __ldg(dYm + static_cast<long long>(b) * ldy + n);
        const float4 w = __ldg(Wm + static_cast<long long>(n) * K4 + k4);
        s.x = fmaf(g, w.x, s.x); s.y = fmaf(g, w.y, s.y); s.z = fmaf(g, w.z, s.z); s.w = fmaf(g, w.w, s.w);
      }
    }
    if (mask_src) {
      const float4 m = ldg_stream(reinterpret_cast<const float4*>(mask_src + model * s_mask + static_cast<long long>(b) * ld_mask) + k4);
      if (mask_mode == PGF_ACT_TANH) {
        s.x *= 1.f - m.x * m.x; s.y *= 1.f - m.y * m.y; s.z *= 1.f - m.z * m.z; s.w *= 1.f - m.w * m.w;
      } else {
        s.x = m.x > 0.f ? s.x : 0.f; s.y = m.y > 0.f ? s.y : 0.f; s.z = m.z > 0.f ? s.z : 0.f; s.w = m.w > 0.f ? s.w : 0.f;
      }
    }
    *(reinterpret_cast<float4*>(dX + model * sdX + static_cast<long long>(b) * ldx) + k4) = s;
  }
}

bool linear_dx_narrow_applies(int B, int N) { return N <= 8 && B >= 512; }

int linear_bwd_dx_narrow(const float* dY, long long ldy, long long sdY, const float* W, long long sW, const float* mask_src,
                         int mask_mode, long long ld_mask, long long s_mask, float* dX, long long ldx, long long sdX, int B, int N, int K,
                         int n_models, cudaStream_t s) {
  const long long total = static_cast<long long>(B) * (K / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (16LL * num_sms() + n_models - 1) / n_models;
  if (blocks > cap) blocks = cap;
  const dim3 grid(static_cast<unsigned>(blocks), n_models);
  if (N <= 2) linear_dx_narrow_kernel<2><<<grid, 256, 0, s>>>(dY, ldy, sdY, W, sW, mask_src, mask_mode, ld_mask, s_mask, dX, ldx, sdX, B, N, K);
  else if (N <= 4) linear_dx_narrow_kernel<4><<<grid, 256, 0, s>>>(dY, ldy, sdY, W, sW, mask_src, mask_mode, ld_mask, s_mask, dX, ldx, sdX, B, N, K);
  else linear_dx_narrow_kernel<8><<<grid, 256, 0, s>>>(dY, ldy, sdY, W, sW, mask_src, mask_mode, ld_mask, s_mask, dX, ldx, sdX, B, N, K);
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_bwd_dx");
  return PGF_OK;
}

bool linear_dw_batch_applies(int B, int N) { return N <= 8 && B >= 512; }

static int linear_dw_batch_slabs(int B, int K, int n_models) {
  const int kctas = (K / 4 + 127) / 128;
  long long slabs = (4LL * num_sms() + static_cast<long long>(kctas) * n_models - 1) / (static_cast<long long>(kctas) * n_models);
  const long long max_slabs = (B + 63) / 64, min_slabs = (B + 1023) / 1024;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < min_slabs) slabs = min_slabs;
  if (slabs < 1) slabs = 1;
  return static_cast<int>(slabs);
}

size_t linear_dw_batch_workspace(int B, int N, int K, int n_models) {
  if (!linear_dw_batch_applies(B, N)) return 0;
  const size_t slabs = static_cast<size_t>(linear_dw_batch_slabs(B, K, n_models));
  return static_cast<size_t>(n_models) * slabs * N * (static_cast<size_t>(K) + 1) * sizeof(float);
}

int linear_bwd_dw_batch(const LinDwArgs& in, int n_models, float* workspace, size_t workspace_bytes, cudaStream_t s) {
  if (workspace_bytes < linear_dw_batch_workspace(in.B, in.N, in.K, n_models)) {
    set_error("pgf_linear_bwd_dw_ex: workspace too small");
    return PGF_ERR_WORKSPACE;
  }
  LinDwBatchArgs a;
  a.dY = in.dY; a.ldy = in.ldy; a.sdY = in.sdY; a.X = in.X; a.ldx = in.ldx; a.sX = in.sX;
  a.B = in.B; a.N = in.N; a.K = in.K;
  a.nslab = linear_dw_batch_slabs(in.B, in.K, n_models);
  a.rows_per_slab = (in.B + a.nslab - 1) / a.nslab;
  a.nslab = (in.B + a.rows_per_slab - 1) / a.rows_per_slab;
  a.partial = workspace;
  a.partial_db = in.db ? workspace + static_cast<size_t>(n_models) * a.nslab * in.N * in.K : nullptr;
  const dim3 grid((in.K / 4 + 127) / 128, a.nslab, n_models);
  const size_t smem = static_cast<size_t>(a.rows_per_slab) * in.N * sizeof(float);
  if (in.N <= 2) linear_dw_batch_kernel<2><<<grid, 128, smem, s>>>(a);
  else if (in.N <= 4) linear_dw_batch_kernel<4><<<grid, 128, smem, s>>>(a);
  else linear_dw_batch_kernel<8><<<grid, 128, smem, s>>>(a);
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_bwd_dw_ex");
  for (int m = 0; m < n_models; ++m) {
    int rc = reduce_partials(a.partial + static_cast<size_t>(m) * a.nslab * in.N * in.K, a.nslab, in.N * in.K, nullptr,
                             in.dW + m * in.sdW, in.accumulate, s);
    if (rc == PGF_OK && in.db)
      rc = reduce_partials(a.partial_db + static_cast<size_t>(m) * a.nslab * in.N, a.nslab, in.N, nullptr, in.db + m * in.sdb,
                           in.accumulate, s);
    if (rc != PGF_OK) return rc;
  }
  return PGF_OK;
}

}  // namespace pgf
