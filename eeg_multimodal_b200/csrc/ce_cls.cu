// Kernel (c): classifier (768->2) + softmax cross-entropy + accuracy, forward and backward fused.
//
// Reference ops replaced: `self.classifier(feature)` (models.py:81) and `cal_loss`
// (base_train.py:59-65 == past_acc.py:71-77: mean F.cross_entropy, argmax, accuracy) plus
// their autograd: dlogits = (softmax - onehot) * scale, dWc, dbc, and the gradient through the
// Tanh that precedes the classifier (dZ2 = (dlogits . Wc) * (1 - h^2)), so the [B,2] logits and
// their gradient never make a separate round trip through HBM.
// One warp owns a row of h [H<=1024]; Wc lives in registers; per-CTA partial sums of
// loss / n_correct / dWc / dbc go to a workspace and are combined by a second, deterministic
// kernel (no float atomics).  Folding the classifier in makes this read h (4*H or 2*H bytes per
// sample): it is HBM-bound instead of launch-bound (roofline for (c)).
#include "pgf_kernels.cuh"

namespace pgf {


template <typename T>
__device__ __forceinline__ float4 ld4(const void* p, long long off);
template <>
__device__ __forceinline__ float4 ld4<float>(const void* p, long long off) {
  return *reinterpret_cast<const float4*>(static_cast<const float*>(p) + off);
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const void* p, long long off) {
  const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p) + off);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void st4(void* p, long long off, const float4& v);
template <>
__device__ __forceinline__ void st4<float>(void* p, long long off, const float4& v) {
  *reinterpret_cast<float4*>(static_cast<float*>(p) + off) = v;
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(void* p, long long off, const float4& v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p) + off) = u;
}

template <int NV, typename HT, typename DT, bool BWD>
__global__ void __launch_bounds__(256) cls_ce_kernel(const CeArgs a) {
  __shared__ float s_red[8][4];
  extern __shared__ float s_dw[];  // [warps][2*H] (BWD only)
  const int model = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const float* Wc = a.Wc + model * a.sWc;
  float4 w0[NV], w1[NV], dw0[NV], dw1[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int col = (lane + 32 * k) << 2;
    const bool ok = col < a.H;
    w0[k] = ok ? *reinterpret_cast<const float4*>(Wc + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    w1[k] = ok ? *reinterpret_cast<const float4*>(Wc + a.H + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    dw0[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    dw1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float b0 = a.bc[model * a.sbc], b1 = a.bc[model * a.sbc + 1];
  const long long* labels = a.labels ? a.labels + model * a.slab : nullptr;  // NULL: logits/pred only
  float loss_sum = 0.f, correct = 0.f, db0 = 0.f, db1 = 0.f;

  for (long long row = static_cast<long long>(blockIdx.x) * nwarps + warp; row < a.B;
       row += static_cast<long long>(gridDim.x) * nwarps) {
    float4 h[NV];
    float z0 = 0.f, z1 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int col = (lane + 32 * k) << 2;
      h[k] = col < a.H ? ld4<HT>(a.h, model * a.sh + row * a.ldh + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      z0 = fmaf(h[k].x, w0[k].x, z0); z0 = fmaf(h[k].y, w0[k].y, z0);
      z0 = fmaf(h[k].z, w0[k].z, z0); z0 = fmaf(h[k].w, w0[k].w, z0);
      z1 = fmaf(h[k].x, w1[k].x, z1); z1 = fmaf(h[k].y, w1[k].y, z1);
      z1 = fmaf(h[k].z, w1[k].z, z1); z1 = fmaf(h[k].w, w1[k].w, z1);
    }
    z0 = warp_sum(z0) + b0;
    z1 = warp_sum(z1) + b1;
    const int label = labels ? static_cast<int>(labels[row]) : 0;
    const float m = fmaxf(z0, z1);
    const float lse = m + logf(expf(z0 - m) + expf(z1 - m));
    const float loss = lse - (label == 0 ? z0 : z1);
    const int pred = z1 > z0 ? 1 : 0;  // torch.argmax: first index on ties
    if (lane == 0) {
      if (a.logits) {
        a.logits[model * a.slogits + row * 2] = z0;
        a.logits[model * a.slogits + row * 2 + 1] = z1;
      }
      if (a.pred) a.pred[model * a.spred + row] = pred;
      loss_sum += loss;
      correct += (pred == label) ? 1.f : 0.f;
    }
    if (BWD) {
      const float g0 = (expf(z0 - lse) - (label == 0 ? 1.f : 0.f)) * a.grad_scale;
      const float g1 = (expf(z1 - lse) - (label == 1 ? 1.f : 0.f)) * a.grad_scale;
      if (lane == 0) { db0 += g0; db1 += g1; }
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int col = (lane + 32 * k) << 2;
        if (col < a.H) {
          dw0[k].x = fmaf(g0, h[k].x, dw0[k].x); dw0[k].y = fmaf(g0, h[k].y, dw0[k].y);
          dw0[k].z = fmaf(g0, h[k].z, dw0[k].z); dw0[k].w = fmaf(g0, h[k].w, dw0[k].w);
          dw1[k].x = fmaf(g1, h[k].x, dw1[k].x); dw1[k].y = fmaf(g1, h[k].y, dw1[k].y);
          dw1[k].z = fmaf(g1, h[k].z, dw1[k].z); dw1[k].w = fmaf(g1, h[k].w, dw1[k].w);
          if (a.dz) {
            float4 d;
            d.x = fmaf(g0, w0[k].x, g1 * w1[k].x);
            d.y = fmaf(g0, w0[k].y, g1 * w1[k].y);
            d.z = fmaf(g0, w0[k].z, g1 * w1[k].z);
            d.w = fmaf(g0, w0[k].w, g1 * w1[k].w);
            if (a.through_tanh) {
              d.x *= 1.f - h[k].x * h[k].x;
              d.y *= 1.f - h[k].y * h[k].y;
              d.z *= 1.f - h[k].z * h[k].z;
              d.w *= 1.f - h[k].w * h[k].w;
            }
            st4<DT>(a.dz, model * a.sdz + row * a.lddz + col, d);
          }
        }
      }
    }
  }
  // ---- CTA-level combine, fixed order
  float* P = a.partial + (static_cast<long long>(model) * gridDim.x + blockIdx.x) * (2 * a.H + 4);
  if (lane == 0) {
    s_red[warp][0] = loss_sum;
    s_red[warp][1] = correct;
    s_red[warp][2] = db0;
    s_red[warp][3] = db1;
  }
  if (BWD) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int col = (lane + 32 * k) << 2;
      if (col < a.H) {
        *reinterpret_cast<float4*>(s_dw + warp * 2 * a.H + col) = dw0[k];
        *reinterpret_cast<float4*>(s_dw + warp * 2 * a.H + a.H + col) = dw1[k];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += s_red[w][threadIdx.x];
    P[threadIdx.x] = s;
  }
  if (BWD) {
    for (int i = threadIdx.x; i < 2 * a.H; i += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < nwarps; ++w) s += s_dw[w * 2 * a.H + i];
      P[4 + i] = s;
    }
  }
}

// stats[model*4 + {0,1,2,3}] = {loss_sum*loss_scale, n_correct, n_correct*loss_scale(acc), B}
__global__ void cls_ce_finalize_kernel(const float* __restrict__ partial, int nctas, int H, int bwd, float loss_scale,
                                       float B, float* __restrict__ stats, float* __restrict__ dWc, long long sdWc,
                                       float* __restrict__ dbc, long long sdbc) {
  // thread i owns output i (4 scalars, then 2*H dWc entries): consecutive threads read consecutive
  // addresses of each CTA's partial row, the loop over CTAs runs in a fixed order (deterministic).
  const int model = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = 2 * H + 4;
  const int n_out = bwd ? stride : 4;
  if (i >= n_out) return;
  const float* P = partial + static_cast<long long>(model) * nctas * stride + i;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = 0;
  for (; c + 4 <= nctas; c += 4) {
    s0 += P[static_cast<long long>(c) * stride];
    s1 += P[static_cast<long long>(c + 1) * stride];
    s2 += P[static_cast<long long>(c + 2) * stride];
    s3 += P[static_cast<long long>(c + 3) * stride];
  }
  for (; c < nctas; ++c) s0 += P[static_cast<long long>(c) * stride];
  const float s = (s0 + s1) + (s2 + s3);
  if (i == 0 && stats) stats[model * 4 + 0] = s * loss_scale;
  if (i == 1 && stats) {
    stats[model * 4 + 1] = s;
    stats[model * 4 + 2] = s * loss_scale;
    stats[model * 4 + 3] = B;
  }
  if (bwd && (i == 2 || i == 3) && dbc) dbc[model * sdbc + (i - 2)] = s;
  if (bwd && i >= 4 && dWc) dWc[model * sdWc + (i - 4)] = s;
}

int cls_ce_ctas(int B, int n_models) {
  int ctas = (num_sms() + n_models - 1) / n_models;
  const int max_ctas = (B + 7) / 8;
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  return ctas;
}

size_t cls_ce_workspace(int B, int H, int n_models) {
  return static_cast<size_t>(n_models) * cls_ce_ctas(B, n_models) * (2 * H + 4) * sizeof(float);
}

template <int NV>
static int launch_ce(const CeArgs& a, int h_dtype, int dz_dtype, bool bwd, int n_models, int ctas, cudaStream_t s) {
  const dim3 grid(ctas, n_models), block(256);
  const size_t smem = bwd ? static_cast<size_t>(8) * 2 * a.H * sizeof(float) : 0;
#define PGF_CE_LAUNCH(HT, DT, BW)                                                                        \
  do {                                                                                                   \
    if (smem > 32 * 1024)                                                                                \
      cudaFuncSetAttribute(cls_ce_kernel<NV, HT, DT, BW>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                           static_cast<int>(smem));                                                      \
    cls_ce_kernel<NV, HT, DT, BW><<<grid, block, smem, s>>>(a);                                          \
  } while (0)
  if (!bwd) {
    if (h_dtype == PGF_DT_F32) PGF_CE_LAUNCH(float, float, false);
    else PGF_CE_LAUNCH(__nv_bfloat16, float, false);
  } else if (h_dtype == PGF_DT_F32) {
    if (dz_dtype == PGF_DT_F32) PGF_CE_LAUNCH(float, float, true);
    else PGF_CE_LAUNCH(float, __nv_bfloat16, true);
  } else {
    if (dz_dtype == PGF_DT_F32) PGF_CE_LAUNCH(__nv_bfloat16, float, true);
    else PGF_CE_LAUNCH(__nv_bfloat16, __nv_bfloat16, true);
  }
#undef PGF_CE_LAUNCH
  PGF_CUDA_LAUNCH_CHECK("pgf_cls_ce");
  return PGF_OK;
}

int cls_ce(const CeArgs& a_in, int h_dtype, int dz_dtype, int bwd, int n_models, float loss_scale, float* stats,
           float* dWc, long long sdWc, float* dbc, long long sdbc, float* workspace, size_t workspace_bytes,
           cudaStream_t s) {
  CeArgs a = a_in;
  if (a.H % 4 != 0 || a.H > 1024) {
    set_error("pgf_cls_ce: hidden width H=%d must be a multiple of 4 and <= 1024", a.H);
    return PGF_ERR_UNSUPPORTED;
  }
  if (workspace_bytes < cls_ce_workspace(a.B, a.H, n_models)) {
    set_error("pgf_cls_ce: workspace too small");
    return PGF_ERR_WORKSPACE;
  }
  a.partial = workspace;
  const int ctas = cls_ce_ctas(a.B, n_models);
  const int nv = (a.H / 4 + 31) / 32;
  int rc;
  if (nv <= 2) rc = launch_ce<2>(a, h_dtype, dz_dtype, bwd != 0, n_models, ctas, s);
  else if (nv <= 4) rc = launch_ce<4>(a, h_dtype, dz_dtype, bwd != 0, n_models, ctas, s);
  else if (nv <= 6) rc = launch_ce<6>(a, h_dtype, dz_dtype, bwd != 0, n_models, ctas, s);
  else rc = launch_ce<8>(a, h_dtype, dz_dtype, bwd != 0, n_models, ctas, s);
  if (rc != PGF_OK) return rc;
  const int n = 2 * a.H + 4;
  const dim3 fgrid((n + 127) / 128, n_models);
  cls_ce_finalize_kernel<<<fgrid, 128, 0, s>>>(workspace, ctas, a.H, bwd, loss_scale, static_cast<float>(a.B), stats, dWc,
                                               sdWc, dbc, sdbc);
  PGF_CUDA_LAUNCH_CHECK("pgf_cls_ce(finalize)");
  return PGF_OK;
}

}  // namespace pgf
