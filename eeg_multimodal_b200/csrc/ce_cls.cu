// Kernel (c): classifier (768->2) + softmax cross-entropy + accuracy, forward and backward fused.
//
// Reference ops replaced: `self.classifier(feature)` (models.py:81) and `cal_loss`
// (base_train.py:59-65 == past_acc.py:71-77: mean F.cross_entropy, argmax, accuracy) plus
// their autograd: dlogits = (softmax - onehot) * scale, dWc, dbc, and the gradient through the
// Tanh that precedes the classifier (dZ2 = (dlogits . Wc) * (1 - h^2)), so the [B,2] logits and
// their gradient never make a separate round trip through HBM.
// One warp owns a row of h [H<=1024]; Wc is staged in shared memory; per-CTA partial sums of
// loss / n_correct / dWc / dbc / colsum(dz) go to a workspace and are combined by a second,
// deterministic kernel (no float atomics).  Folding the classifier in makes this read h (4*H or 2*H bytes per
// sample): it is HBM-bound instead of launch-bound (roofline for (c)).
#include "pgf_kernels.cuh"

namespace pgf {


template <typename T>
__device__ __forceinline__ float4 ld4(const void* p, long long off);
template <>
__device__ __forceinline__ float4 ld4<float>(const void* p, long long off) {
  return ldg_stream(reinterpret_cast<const float4*>(static_cast<const float*>(p) + off));
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const void* p, long long off) {
  const uint2 u = ldg_stream_u2(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p) + off));
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void st4(void* p, long long off, const float4& v);
template <>
__device__ __forceinline__ void st4<float>(void* p, long long off, const float4& v) {
  stg_stream(reinterpret_cast<float4*>(static_cast<float*>(p) + off), v);
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(void* p, long long off, const float4& v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  stg_stream_u2(reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p) + off), u);
}

// MODE 0: forward only (eval).  MODE 1: + dz (pass 1 of the reference step: only the dX chain is
// needed).  MODE 2: + dWc, dbc and the column sums of dz (= the bias gradient of fc_layers.2), pass 2.
// Each warp streams its rows through a private 3-stage ring of shared-memory row buffers filled by bulk async
// copies (cp.async.bulk + mbarrier complete_tx, issued by lane 0 two rows ahead): register prefetching is
// defeated by scoreboard aliasing (ncu: 25-35 % of the stall samples on the first FFMA of a row), and the ring
// costs no registers, which MODE 2 needs for its gradient accumulators.  Wc is staged in shared memory.
constexpr int CE_THREADS = 256;
constexpr int CE_STAGES = 3;

__device__ __forceinline__ uint32_t ce_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ce_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
template <typename T>
__device__ __forceinline__ float4 lds4(const unsigned char* row, int j);
template <>
__device__ __forceinline__ float4 lds4<float>(const unsigned char* row, int j) {
  return reinterpret_cast<const float4*>(row)[j];
}
template <>
__device__ __forceinline__ float4 lds4<__nv_bfloat16>(const unsigned char* row, int j) {
  const uint2 u = reinterpret_cast<const uint2*>(row)[j];
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

// dynamic shared memory: [Wc 2*H floats][ring: warps x CE_STAGES x row bytes, reused for the dWc combine][MODE 2: warps x H colsum]
__host__ __device__ inline size_t ce_ring_bytes(int H, size_t esz, int mode) {
  const size_t ring = static_cast<size_t>(CE_THREADS / 32) * CE_STAGES * H * esz;
  const size_t comb = mode == 2 ? static_cast<size_t>(CE_THREADS / 32) * 2 * H * sizeof(float) : 0;
  return ring > comb ? ring : comb;
}

template <int NV, typename HT, typename DT, int MODE>
__global__ void __launch_bounds__(CE_THREADS, 2) cls_ce_kernel(const CeArgs a) {
  __shared__ float s_red[CE_THREADS / 32][4];
  __shared__ __align__(8) unsigned long long s_full[CE_THREADS / 32][CE_STAGES];
  extern __shared__ float4 s_dyn4[];
  const int model = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = CE_THREADS / 32;
  const int nvec = a.H >> 2;
  float4* s_w = s_dyn4;
  unsigned char* s_ring = reinterpret_cast<unsigned char*>(s_dyn4 + 2 * nvec);
  const uint32_t row_bytes = static_cast<uint32_t>(a.H) * static_cast<uint32_t>(sizeof(HT));
  unsigned char* my_ring = s_ring + static_cast<size_t>(warp) * CE_STAGES * row_bytes;
  float* s_cs = reinterpret_cast<float*>(s_ring + ce_ring_bytes(a.H, sizeof(HT), MODE));  // [warps][H] (MODE 2)
  {
    const float4* Wc = reinterpret_cast<const float4*>(a.Wc + model * a.sWc);
    for (int i = threadIdx.x; i < 2 * nvec; i += CE_THREADS) s_w[i] = Wc[i];
    if (lane == 0) {
#pragma unroll
      for (int st = 0; st < CE_STAGES; ++st) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ce_smem_u32(&s_full[warp][st])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
  }
  griddep_wait();     // Wc above never comes from the preceding kernel of a step; the activations below do
  griddep_launch();
  const long long cursor = (a.st && a.gather) ? a.st->cursor : 0;
  // MODE 2 accumulators: dWc rows in registers, the dz column sums in the warp's own shared-memory row
  float4 dw0[MODE == 2 ? NV : 1], dw1[MODE == 2 ? NV : 1];
  float* s_dsum = s_cs + warp * a.H;
  if (MODE == 2) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      dw0[k] = dw1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int col = (lane + 32 * k) << 2;
      if (col < a.H) *reinterpret_cast<float4*>(s_dsum + col) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // MODE 0/1 have the registers to keep this lane's slice of Wc resident (MODE 2 spends them on the dWc accumulators
  // and re-reads Wc from shared memory per row)
  constexpr bool WREG = MODE != 2;
  float4 wr0[WREG ? NV : 1], wr1[WREG ? NV : 1];
  if (WREG) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int j = lane + 32 * k;
      const bool ok = (j << 2) < a.H;
      wr0[k] = ok ? s_w[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      wr1[k] = ok ? s_w[nvec + j] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float b0 = a.bc[model * a.sbc], b1 = a.bc[model * a.sbc + 1];
  const long long* labels = a.labels ? a.labels + model * a.slab : nullptr;  // NULL: logits/pred only
  float loss_sum = 0.f, correct = 0.f, db0 = 0.f, db1 = 0.f;
  const long long row_step = static_cast<long long>(gridDim.x) * nwarps;
  const long long row_first = static_cast<long long>(blockIdx.x) * nwarps + warp;
  const HT* hsrc = static_cast<const HT*>(a.h) + model * a.sh;

  auto issue = [&](int it) {  // lane 0: request row `it` of this warp
    const long long r = row_first + static_cast<long long>(it) * row_step;
    if (r >= a.B) return;
    const int st = it % CE_STAGES;
    const uint32_t bar = ce_smem_u32(&s_full[warp][st]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     ce_smem_u32(my_ring + static_cast<size_t>(st) * row_bytes)),
                 "l"(hsrc + r * a.ldh), "r"(row_bytes), "r"(bar)
                 : "memory");
  };
  if (lane == 0) {
#pragma unroll
    for (int it = 0; it < CE_STAGES - 1; ++it) issue(it);
  }

  int it = 0;
  for (long long row = row_first; row < a.B; row += row_step, ++it) {
    const int st = it % CE_STAGES;
    ce_mbar_wait(ce_smem_u32(&s_full[warp][st]), static_cast<uint32_t>(it / CE_STAGES) & 1u);
    float4 h[NV];
    const unsigned char* srow = my_ring + static_cast<size_t>(st) * row_bytes;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int j = lane + 32 * k;
      h[k] = (j << 2) < a.H ? lds4<HT>(srow, j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();                                  // every lane holds its part of the row: the previous stage is free
    if (lane == 0) issue(it + CE_STAGES - 1);
    float z0 = 0.f, z1 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int j = lane + 32 * k;
      if ((j << 2) < a.H) {
        const float4 w0 = WREG ? wr0[k] : s_w[j], w1 = WREG ? wr1[k] : s_w[nvec + j];
        z0 = fmaf(h[k].x, w0.x, z0); z0 = fmaf(h[k].y, w0.y, z0);
        z0 = fmaf(h[k].z, w0.z, z0); z0 = fmaf(h[k].w, w0.w, z0);
        z1 = fmaf(h[k].x, w1.x, z1); z1 = fmaf(h[k].y, w1.y, z1);
        z1 = fmaf(h[k].z, w1.z, z1); z1 = fmaf(h[k].w, w1.w, z1);
      }
    }
    z0 = warp_sum(z0) + b0;
    z1 = warp_sum(z1) + b1;
    const long long lrow = a.gather ? (a.src_rows ? a.src_rows[cursor + row] : cursor + row) : row;
    const int label = labels ? static_cast<int>(labels[lrow]) : 0;
    const float m = fmaxf(z0, z1);
    const float lse = m + logf(expf(z0 - m) + expf(z1 - m));
    // a label outside {0,1} (F.cross_entropy raises on it) poisons the loss statistic with NaN instead of being scored silently
    const float loss = (label == 0 || label == 1) ? lse - (label == 0 ? z0 : z1) : __int_as_float(0x7fc00000);
    const int pred = z1 > z0 ? 1 : 0;  // torch.argmax: first index on ties
    if (lane == 0) {
      if (a.logits) *reinterpret_cast<float2*>(a.logits + model * a.slogits + row * 2) = make_float2(z0, z1);
      if (a.pred) a.pred[model * a.spred + row] = pred;
      loss_sum += loss;
      correct += (pred == label) ? 1.f : 0.f;
    }
    if (MODE >= 1) {
      const float g0 = (expf(z0 - lse) - (label == 0 ? 1.f : 0.f)) * a.grad_scale;
      const float g1 = (expf(z1 - lse) - (label == 1 ? 1.f : 0.f)) * a.grad_scale;
      if (MODE == 2 && lane == 0) { db0 += g0; db1 += g1; }
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int j = lane + 32 * k;
        if ((j << 2) < a.H) {
          if (MODE == 2) {
            dw0[k].x = fmaf(g0, h[k].x, dw0[k].x); dw0[k].y = fmaf(g0, h[k].y, dw0[k].y);
            dw0[k].z = fmaf(g0, h[k].z, dw0[k].z); dw0[k].w = fmaf(g0, h[k].w, dw0[k].w);
            dw1[k].x = fmaf(g1, h[k].x, dw1[k].x); dw1[k].y = fmaf(g1, h[k].y, dw1[k].y);
            dw1[k].z = fmaf(g1, h[k].z, dw1[k].z); dw1[k].w = fmaf(g1, h[k].w, dw1[k].w);
          }
          if (a.dz) {
            const float4 w0 = WREG ? wr0[k] : s_w[j], w1 = WREG ? wr1[k] : s_w[nvec + j];
            float4 d;
            d.x = fmaf(g0, w0.x, g1 * w1.x);
            d.y = fmaf(g0, w0.y, g1 * w1.y);
            d.z = fmaf(g0, w0.z, g1 * w1.z);
            d.w = fmaf(g0, w0.w, g1 * w1.w);
            if (a.through_tanh) {
              d.x *= 1.f - h[k].x * h[k].x;
              d.y *= 1.f - h[k].y * h[k].y;
              d.z *= 1.f - h[k].z * h[k].z;
              d.w *= 1.f - h[k].w * h[k].w;
            }
            if (MODE == 2) {
              float4 acc = *reinterpret_cast<float4*>(s_dsum + (j << 2));
              acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
              *reinterpret_cast<float4*>(s_dsum + (j << 2)) = acc;
            }
            st4<DT>(a.dz, model * a.sdz + row * a.lddz + (j << 2), d);
          }
        }
      }
    }
  }
  // ---- CTA-level combine, fixed order.  Partial row of a CTA: [4 scalars | dWc 2H | dz column sums H]
  const int pstride = 3 * a.H + 4;
  float* P = a.partial + (static_cast<long long>(model) * gridDim.x + blockIdx.x) * pstride;
  if (lane == 0) {
    s_red[warp][0] = loss_sum;
    s_red[warp][1] = correct;
    s_red[warp][2] = db0;
    s_red[warp][3] = db1;
  }
  __syncthreads();  // every warp is done with its ring: the ring space now stages the per-warp dWc rows
  float* s_dw = reinterpret_cast<float*>(s_ring);  // [warps][2*H]
  if (MODE == 2) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int col = (lane + 32 * k) << 2;
      if (col < a.H) {
        *reinterpret_cast<float4*>(s_dw + warp * 2 * a.H + col) = dw0[k];
        *reinterpret_cast<float4*>(s_dw + warp * 2 * a.H + a.H + col) = dw1[k];
      }
    }
    __syncthreads();
  }
  const bool direct = a.direct != 0;   // this CTA is the model's only one: its sums ARE the outputs (what the finalize kernel
                                       // would compute from a single partial row, bit for bit)
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += s_red[w][threadIdx.x];
    if (!direct) {
      P[threadIdx.x] = s;
    } else {
      const int i = threadIdx.x;
      if (i == 0 && a.stats) a.stats[model * 4 + 0] = s * a.loss_scale;
      if (i == 1 && a.stats) {
        a.stats[model * 4 + 1] = s;
        a.stats[model * 4 + 2] = s * a.loss_scale;
        a.stats[model * 4 + 3] = static_cast<float>(a.B);
      }
      if (MODE == 2 && i >= 2 && a.dbc) a.dbc[model * a.sdbc + (i - 2)] = s;
      if (MODE == 2 && i >= 2 && a.adam_mb) {   // fused model_optimizer.step() on classifier.bias
        const AdamCoef c = adam_coef_at(a.adam_c, a.st, 1);
        const long long o = model * a.sbc + (i - 2);
        float p = const_cast<float*>(a.bc)[o], m = a.adam_mb[o], v = a.adam_vb[o];
        adam_update(p, m, v, s, c);
        const_cast<float*>(a.bc)[o] = p; a.adam_mb[o] = m; a.adam_vb[o] = v;
      }
    }
  }
  if (MODE == 2) {
    for (int i = threadIdx.x; i < 3 * a.H; i += CE_THREADS) {
      float s = 0.f;
      if (i < 2 * a.H) {
        for (int w = 0; w < nwarps; ++w) s += s_dw[w * 2 * a.H + i];
      } else {
        for (int w = 0; w < nwarps; ++w) s += s_cs[w * a.H + (i - 2 * a.H)];
      }
      if (!direct) P[4 + i] = s;
      else if (i < 2 * a.H) {
        if (a.dWc) a.dWc[model * a.sdWc + i] = s;
        if (a.adam_m) {   // fused model_optimizer.step() on classifier.weight (every row read its copy in shared memory)
          const AdamCoef c = adam_coef_at(a.adam_c, a.st, 1);
          const long long o = model * a.sWc + i;
          float p = const_cast<float*>(a.Wc)[o], m = a.adam_m[o], v = a.adam_v[o];
          adam_update(p, m, v, s, c);
          const_cast<float*>(a.Wc)[o] = p; a.adam_m[o] = m; a.adam_v[o] = v;
        }
      }
      else if (a.dzsum) a.dzsum[model * a.sdzsum + (i - 2 * a.H)] = s;
    }
  }
}

// stats[model*4 + {0,1,2,3}] = {loss_sum*loss_scale, n_correct, n_correct*loss_scale(acc), B}
// Output i of a model (4 scalars, 2*H dWc entries, H dz column sums) is the sum over the CTAs' partial rows.
// A CTA owns 32 consecutive outputs (coalesced 128-byte reads of every partial row); its 8 warps take
// interleaved rows with 4 loads in flight each and combine in a fixed order: deterministic, and the
// dependent-load chain is nctas/32 long instead of nctas/4.
__global__ void __launch_bounds__(256) cls_ce_finalize_kernel(const float* __restrict__ partial, int nctas, int H, int mode,
                                                              float loss_scale, float B, float* __restrict__ stats,
                                                              float* __restrict__ dWc, long long sdWc, float* __restrict__ dbc,
                                                              long long sdbc, float* __restrict__ dzsum, long long sdzsum) {
  __shared__ float s_part[8][32];
  const int model = blockIdx.y;
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const int stride = 3 * H + 4;
  const int n_out = mode == 2 ? stride : 4;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < n_out) {
    const float* P = partial + static_cast<long long>(model) * nctas * stride + i;
    int c = rg;
    for (; c + 24 < nctas; c += 32) {
      s0 += P[static_cast<long long>(c) * stride];
      s1 += P[static_cast<long long>(c + 8) * stride];
      s2 += P[static_cast<long long>(c + 16) * stride];
      s3 += P[static_cast<long long>(c + 24) * stride];
    }
    for (; c < nctas; c += 8) s0 += P[static_cast<long long>(c) * stride];
  }
  s_part[rg][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (rg != 0 || i >= n_out) return;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += s_part[w][lane];
  if (i == 0 && stats) stats[model * 4 + 0] = s * loss_scale;
  if (i == 1 && stats) {
    stats[model * 4 + 1] = s;
    stats[model * 4 + 2] = s * loss_scale;
    stats[model * 4 + 3] = B;
  }
  if (mode == 2) {
    if ((i == 2 || i == 3) && dbc) dbc[model * sdbc + (i - 2)] = s;
    if (i >= 4 && i < 4 + 2 * H && dWc) dWc[model * sdWc + (i - 4)] = s;
    if (i >= 4 + 2 * H && dzsum) dzsum[model * sdzsum + (i - 4 - 2 * H)] = s;
  }
}

int cls_ce_ctas(int B, int n_models) {
  int ctas = (2 * num_sms()) / n_models;  // one resident wave (2 CTAs per SM) over the whole grouped launch: no tail
  const int max_ctas = (B + 7) / 8;
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  return ctas;
}

size_t cls_ce_workspace(int B, int H, int n_models) {
  return static_cast<size_t>(n_models) * cls_ce_ctas(B, n_models) * (3 * H + 4) * sizeof(float);
}

template <int NV>
static int launch_ce(const CeArgs& a, int h_dtype, int dz_dtype, int mode, int n_models, int ctas, cudaStream_t s) {
  const dim3 grid(ctas, n_models), block(CE_THREADS);
  const size_t esz = h_dtype == PGF_DT_F32 ? 4 : 2;
  const size_t smem = static_cast<size_t>(2) * a.H * sizeof(float) + ce_ring_bytes(a.H, esz, mode) +
                      (mode == 2 ? static_cast<size_t>(CE_THREADS / 32) * a.H * sizeof(float) : 0);
#define PGF_CE_LAUNCH(HT, DT, MD)                                                                        \
  do {                                                                                                   \
    ensure_dynamic_smem(reinterpret_cast<const void*>(cls_ce_kernel<NV, HT, DT, MD>), smem);              \
    launch(cls_ce_kernel<NV, HT, DT, MD>, grid, block, smem, s, a);                                      \
  } while (0)
#define PGF_CE_MODES(HT, DT)                         \
  do {                                               \
    if (mode == 2) PGF_CE_LAUNCH(HT, DT, 2);         \
    else if (mode == 1) PGF_CE_LAUNCH(HT, DT, 1);    \
    else PGF_CE_LAUNCH(HT, DT, 0);                   \
  } while (0)
  if (h_dtype == PGF_DT_F32) {
    if (dz_dtype == PGF_DT_F32) PGF_CE_MODES(float, float);
    else PGF_CE_MODES(float, __nv_bfloat16);
  } else {
    if (dz_dtype == PGF_DT_F32) PGF_CE_MODES(__nv_bfloat16, float);
    else PGF_CE_MODES(__nv_bfloat16, __nv_bfloat16);
  }
#undef PGF_CE_MODES
#undef PGF_CE_LAUNCH
  PGF_CUDA_LAUNCH_CHECK("pgf_cls_ce");
  return PGF_OK;
}

int cls_ce(const CeArgs& a_in, int h_dtype, int dz_dtype, int bwd, int n_models, float loss_scale, float* stats,
           float* dWc, long long sdWc, float* dbc, long long sdbc, float* dzsum, long long sdzsum, float* workspace,
           size_t workspace_bytes, cudaStream_t s) {
  CeArgs a = a_in;
  if (a.H % 4 != 0 || a.H > 1024) {
    set_error("pgf_cls_ce: hidden width H=%d must be a multiple of 4 and <= 1024", a.H);
    return PGF_ERR_UNSUPPORTED;
  }
  if (h_dtype != PGF_DT_F32 && ((a.H % 8) || (a.ldh % 8) || (a.sh % 8))) {
    set_error("pgf_cls_ce: bf16 activations need H, ldh and the model stride to be multiples of 8 (16-byte bulk copies)");
    return PGF_ERR_ARG;
  }
  if (workspace_bytes < cls_ce_workspace(a.B, a.H, n_models)) {
    set_error("pgf_cls_ce: workspace too small");
    return PGF_ERR_WORKSPACE;
  }
  a.partial = workspace;
  // the weight-side gradients are formed only when the caller asks for one of them
  const int mode = !bwd ? 0 : ((dWc || dbc || dzsum) ? 2 : 1);
  const int ctas = cls_ce_ctas(a.B, n_models);
  a.direct = ctas == 1;
  if ((a.adam_m || a.adam_mb) && !(a.direct && mode == 2)) {
    set_error("pgf_cls_ce: the fused classifier Adam needs the one-CTA-per-model pass-2 launch (B <= 8)");
    return PGF_ERR_UNSUPPORTED;
  }
  a.loss_scale = loss_scale;
  a.stats = stats; a.dWc = dWc; a.sdWc = sdWc; a.dbc = dbc; a.sdbc = sdbc; a.dzsum = dzsum; a.sdzsum = sdzsum;
  const int nv = (a.H / 4 + 31) / 32;
  int rc;
  if (nv <= 2) rc = launch_ce<2>(a, h_dtype, dz_dtype, mode, n_models, ctas, s);
  else if (nv <= 4) rc = launch_ce<4>(a, h_dtype, dz_dtype, mode, n_models, ctas, s);
  else if (nv <= 6) rc = launch_ce<6>(a, h_dtype, dz_dtype, mode, n_models, ctas, s);
  else rc = launch_ce<8>(a, h_dtype, dz_dtype, mode, n_models, ctas, s);
  if (rc != PGF_OK) return rc;
  if (a.direct) return PGF_OK;
  const int n = mode == 2 ? 3 * a.H + 4 : 4;
  const dim3 fgrid((n + 31) / 32, n_models);
  cls_ce_finalize_kernel<<<fgrid, 256, 0, s>>>(workspace, ctas, a.H, mode, loss_scale, static_cast<float>(a.B), stats, dWc,
                                               sdWc, dbc, sdbc, dzsum, sdzsum);
  PGF_CUDA_LAUNCH_CHECK("pgf_cls_ce(finalize)");
  return PGF_OK;
}

}  // namespace pgf
