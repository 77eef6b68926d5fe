// Kernel (a): fused concat -> row min-max normalise -> Laplace perturbation -> Gumbel gate.
//
// Reference ops replaced (python/src/custom_models/models.py, == past_acc.py:120-136):
//   :69  torch.cat of the feature blocks
//   :70-72 row min / row max / (x-min)/(max-min)          (no epsilon guard, NaN on constant row)
//   :74  Laplace(0,1).sample on the HOST + .to(device)    -> Philox in-kernel, or injected tensor
//   :76  feature + noise * eps_hat
//   :77-79 gumbel_softmax over the stacked (w, 1-w) planes and (feature*mask).sum(0)
// One warp owns one row: the whole row (<= 4096 floats) lives in registers, min/max are
// warp-shuffle reductions, every global access is a coalesced 128-bit load/store, and the
// per-column coefficients (eps_hat, w) are staged once per CTA in shared memory.
//
// HBM traffic per row (Philox mode): read 4*D, write 4*D (fp32 out) or 2*D (bf16 out).
#include <stdlib.h>

#include "pgf_kernels.cuh"
#include "philox.cuh"

namespace pgf {


template <typename OutT>
__device__ __forceinline__ void store_out4(void* out, long long off, const float4& v);
template <>
__device__ __forceinline__ void store_out4<float>(void* out, long long off, const float4& v) {
  stg_stream(reinterpret_cast<float4*>(static_cast<float*>(out) + off), v);
}
template <>
__device__ __forceinline__ void store_out4<__nv_bfloat16>(void* out, long long off, const float4& v) {
  uint2 p;
  p.x = pack_bf16x2(v.x, v.y);
  p.y = pack_bf16x2(v.z, v.w);
  stg_stream_u2(reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + off), p);
}

// faithful two-plane softmax gate on one element (injected mode): returns gated value and index
__device__ __forceinline__ float gate_faithful(float f, float w, float g0, float g1, float tau, int hard, int& idx) {
  const float z0 = __fdiv_rn(__fadd_rn(w, g0), tau);
  const float z1 = __fdiv_rn(__fadd_rn(__fsub_rn(1.0f, w), g1), tau);
  const float m = fmaxf(z0, z1);
  const float e0 = expf(z0 - m), e1 = expf(z1 - m);
  const float s = __fadd_rn(e0, e1);
  const float y0 = __fdiv_rn(e0, s), y1 = __fdiv_rn(e1, s);
  idx = (y1 > y0) ? 1 : 0;  // argmax over dim 0, first index on ties
  float m0 = y0, m1 = y1;
  if (hard) {  // y_hard - y_soft.detach() + y_soft
    m0 = __fadd_rn(__fsub_rn(idx == 0 ? 1.0f : 0.0f, y0), y0);
    m1 = __fadd_rn(__fsub_rn(idx == 1 ? 1.0f : 0.0f, y1), y1);
  }
  return __fadd_rn(__fmul_rn(f, m0), __fmul_rn(f, m1));
}

constexpr int FWD_THREADS = 128;  // 4 warps share one row: <= 8 float4 per lane, ~64 registers, 32 warps/SM

template <int NV, int NOISE, typename OutT, bool WANT_GATE, bool CONSTKEYS>
__global__ void __launch_bounds__(FWD_THREADS, 6) perturb_gate_fwd_kernel(const PerturbFwdArgs a) {
  extern __shared__ float4 smem4[];
  __shared__ float s_part[2][FWD_THREADS / 32][3];  // [parity][warp]{min, max, nan-probe}
  float4* s_eps = smem4;                 // [D/4]  eps_hat (Philox mode: pre-multiplied by -ln2)
  float4* s_w = smem4 + (a.D >> 2);      // [D/4]  (only when the gate is evaluated)
  const int model = blockIdx.y;
  const int nvec = a.D >> 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  griddep_wait();     // the coefficient rows (and, in a step chain, DP) come from the preceding kernel
  griddep_launch();
  const unsigned int offset0 = a.offset + (a.st ? static_cast<unsigned int>(a.st->noise_offset) : 0u);
  const long long cursor = (a.st && a.gather) ? a.st->cursor : 0;
  const int n_rep = a.n_rep > 1 ? a.n_rep : 1;
  if (NOISE != PGF_NOISE_NONE) {
    const float4* ge = reinterpret_cast<const float4*>(a.eps_hat + model * a.s_coef);
    const float4* gw = reinterpret_cast<const float4*>(a.w + model * a.s_coef);
    for (int i = tid; i < nvec; i += FWD_THREADS) {
      float4 e = ge[i];
      if (NOISE == PGF_NOISE_PHILOX) {  // fold -ln(2) of  -ln(v) = -ln2 * lg2(v)  into the coefficient
        e.x *= -0.69314718055994531f; e.y *= -0.69314718055994531f;
        e.z *= -0.69314718055994531f; e.w *= -0.69314718055994531f;
      }
      s_eps[i] = e;
      if (WANT_GATE) s_w[i] = gw[i];
    }
    __syncthreads();
  }
  const int d01 = a.d[0] + a.d[1];
  const float* x0 = a.x[0] + model * a.sx[0];
  const float* x1 = a.d[1] ? a.x[1] + model * a.sx[1] : nullptr;
  const float* x2 = a.d[2] ? a.x[2] + model * a.sx[2] : nullptr;
  const unsigned long long seed = a.model_seeds ? a.model_seeds[model] : a.seed + static_cast<unsigned long long>(model) * a.seed_step;
  const unsigned int k0 = static_cast<unsigned int>(seed), k1 = static_cast<unsigned int>(seed >> 32);
  const long long BD = static_cast<long long>(a.B) * n_rep * a.D;
  const float* lap = a.lap ? a.lap + model * BD : nullptr;
  const float* gum = a.gum ? a.gum + model * 2 * BD : nullptr;
  unsigned char* gate_idx = a.gate_idx ? a.gate_idx + model * BD : nullptr;

  // fused concat (models.py:69): which block a lane's k-th float4 comes from does not depend on the
  // row, so the (base pointer, row stride) pair is selected once, branch-free, outside the row loop
  const float* pk[NV];
  long long ldk[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int col = (tid + FWD_THREADS * k) << 2;
    const bool in1 = col >= a.d[0], in2 = col >= d01;
    const float* base = in2 ? x2 : (in1 ? x1 : x0);
    const int off = in2 ? col - d01 : (in1 ? col - a.d[0] : col);
    pk[k] = base + off;
    ldk[k] = in2 ? a.ld[2] : (in1 ? a.ld[1] : a.ld[0]);
  }

  int it = 0;
  const long long vrows = static_cast<long long>(a.B) * n_rep;
  for (long long row = blockIdx.x; row < vrows; row += gridDim.x, ++it) {
    // virtual row = (repetition, batch row); the batch row may be gathered from a resident dataset
    const int rep = static_cast<int>(row / a.B);
    const long long brow = row - static_cast<long long>(rep) * a.B;
    const long long srow = a.gather ? (a.src_rows ? a.src_rows[cursor + brow] : cursor + brow) : brow;
    const unsigned int offs = offset0 + static_cast<unsigned int>(rep);
    float4 v[NV];
    float mn = INFINITY, mx = -INFINITY, probe = 0.f;
    // all loads of the row are issued before first use
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int col = (tid + FWD_THREADS * k) << 2;
      if (col < a.D) v[k] = ldg_stream(reinterpret_cast<const float4*>(pk[k] + srow * ldk[k]));
    }
    // ---- row min / max (models.py:70-71).  torch.min/max propagate NaN while fminf/fmaxf drop it, so
    // a running sum is carried as a NaN probe (NaN in -> NaN out; an inf/-inf mix also gives the NaN
    // row the reference produces).
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int col = (tid + FWD_THREADS * k) << 2;
      if (col < a.D) {
        mn = fminf(fminf(mn, v[k].x), fminf(v[k].y, fminf(v[k].z, v[k].w)));
        mx = fmaxf(fmaxf(mx, v[k].x), fmaxf(v[k].y, fmaxf(v[k].z, v[k].w)));
        probe += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      }
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    probe = warp_sum(probe);
    float(*part)[3] = s_part[it & 1];
    if (lane == 0) {
      part[warp][0] = mn;
      part[warp][1] = mx;
      part[warp][2] = probe;
    }
    __syncthreads();  // one barrier per row: the partial slots are double-buffered on the row parity
#pragma unroll
    for (int w = 0; w < FWD_THREADS / 32; ++w) {
      mn = fminf(mn, part[w][0]);
      mx = fmaxf(mx, part[w][1]);
    }
    probe = (part[0][2] + part[1][2]) + (part[2][2] + part[3][2]);
    if (probe != probe) mn = mx = __int_as_float(0x7fc00000);
    const float range = __fsub_rn(mx, mn);
    const float inv_range = __frcp_rn(range);
    if (tid == 0) {
      if (a.row_min) a.row_min[model * vrows + row] = mn;
      if (a.row_max) a.row_max[model * vrows + row] = mx;
    }
    const unsigned int grow = static_cast<unsigned int>(a.row0 + static_cast<unsigned long long>(brow));
    char* outp = static_cast<char*>(a.out) + model * a.s_out * static_cast<long long>(sizeof(OutT));

#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int j = tid + FWD_THREADS * k;
      const int col = j << 2;
      if (col < a.D) {
        float f[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
        // ---- normalise (models.py:72).  Parity modes use the exact IEEE division the
        // reference does; Philox mode multiplies by the reciprocal (<= 1 ulp apart).
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (NOISE == PGF_NOISE_PHILOX)
            f[e] = __fmul_rn(__fsub_rn(f[e], mn), inv_range);
          else
            f[e] = __fdiv_rn(__fsub_rn(f[e], mn), range);
        }
        int idx[4] = {0, 0, 0, 0};
        if (NOISE == PGF_NOISE_INJECTED) {
          const float4 e4 = s_eps[j];
          const float eh[4] = {e4.x, e4.y, e4.z, e4.w};
          const float4 l4 = ldg_stream(reinterpret_cast<const float4*>(lap + row * a.D + col));
          const float lp[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) f[e] = __fadd_rn(f[e], __fmul_rn(lp[e], eh[e]));  // :76
          if (WANT_GATE) {                                                               // :77-79
            const float4 w4 = s_w[j];
            const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
            const float4 g0 = ldg_stream(reinterpret_cast<const float4*>(gum + row * a.D + col));
            const float4 g1 = ldg_stream(reinterpret_cast<const float4*>(gum + BD + row * a.D + col));
            const float ga[4] = {g0.x, g0.y, g0.z, g0.w}, gb[4] = {g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) f[e] = gate_faithful(f[e], ww[e], ga[e], gb[e], a.tau, a.hard, idx[e]);
          }
        } else if (NOISE == PGF_NOISE_PHILOX) {
          const float4 e4 = s_eps[j];
          const float eh[4] = {e4.x, e4.y, e4.z, e4.w};  // = -ln2 * eps_hat
          const uint4 r = CONSTKEYS ? philox4x32_10_rk(static_cast<unsigned int>(j), grow, PGF_STREAM_LAPLACE, offs, a.rk)
                                    : philox4x32_10(static_cast<unsigned int>(j), grow, PGF_STREAM_LAPLACE, offs, k0, k1);
          const unsigned int rb[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) f[e] = perturb_fma(f[e], rb[e], eh[e]);
          if (WANT_GATE) {
            // The two mask planes sum to one (hard: exactly, soft: within 1 ulp), so the gated
            // value IS the perturbed value; only the gate index is a real output.
            const float4 w4 = s_w[j];
            const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
            const uint4 q0 = philox4x32_10(static_cast<unsigned int>(j), grow, PGF_STREAM_GUMBEL0, offs, k0, k1);
            const uint4 q1 = philox4x32_10(static_cast<unsigned int>(j), grow, PGF_STREAM_GUMBEL1, offs, k0, k1);
            const unsigned int qa[4] = {q0.x, q0.y, q0.z, q0.w}, qb[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float z0 = (ww[e] + gumbel_from_bits(qa[e])) * a.inv_tau;
              const float z1 = ((1.0f - ww[e]) + gumbel_from_bits(qb[e])) * a.inv_tau;
              idx[e] = z1 > z0 ? 1 : 0;
            }
          }
        }
        store_out4<OutT>(outp, row * a.ld_out + col, make_float4(f[0], f[1], f[2], f[3]));
        if (WANT_GATE && gate_idx) {
          const unsigned int packed = idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24);
          *reinterpret_cast<unsigned int*>(gate_idx + row * a.D + col) = packed;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Large-batch production variant (Philox / non-private, one model per launch): the CTA's rows arrive
// through a ring of shared-memory row buffers filled by 1-D bulk async copies (TMA engine,
// cp.async.bulk + mbarrier complete_tx), RING_STAGES-1 rows ahead of the arithmetic.  Register-side
// prefetching does not work here: the consumer's wait is a scoreboard wait, and the scoreboard of the
// row being consumed also covers the just-issued loads of the next row (ncu: 30 % of all stall samples on
// the first use of the row).  mbarrier completion has no such aliasing, costs no registers, and one elected
// thread issues the copies, so the other 127 never compute a global address.
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ring_mbar_wait(uint32_t bar, uint32_t parity) {
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

template <int NV, int NOISE, typename OutT, bool CONSTKEYS, int RING_STAGES>
__global__ void __launch_bounds__(FWD_THREADS, 5) perturb_fwd_ring_kernel(const PerturbFwdArgs a) {
  extern __shared__ float4 smem4[];
  __shared__ float s_part[2][FWD_THREADS / 32][3];
  __shared__ __align__(8) unsigned long long s_full[RING_STAGES];
  const int nvec = a.D >> 2;
  float4* s_eps = smem4;                                   // [D/4]  -ln2 * eps_hat
  float4* s_rows = smem4 + (NOISE == PGF_NOISE_NONE ? 0 : nvec);  // [RING_STAGES][D/4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < RING_STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&s_full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (NOISE != PGF_NOISE_NONE) {
    const float4* ge = reinterpret_cast<const float4*>(a.eps_hat);
    for (int i = tid; i < nvec; i += FWD_THREADS) {
      float4 e = ge[i];
      e.x *= -0.69314718055994531f; e.y *= -0.69314718055994531f;
      e.z *= -0.69314718055994531f; e.w *= -0.69314718055994531f;
      s_eps[i] = e;
    }
  }
  __syncthreads();
  const unsigned int k0 = static_cast<unsigned int>(a.seed), k1 = static_cast<unsigned int>(a.seed >> 32);
  const uint32_t row_bytes = static_cast<uint32_t>(a.D) * 4u;

  // producer (thread 0): one bulk copy per feature block = the fused concat (models.py:69)
  auto issue = [&](int it) {
    const long long row = static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x;
    if (row >= a.B) return;
    const int st = it % RING_STAGES;
    const uint32_t bar = smem_addr_u32(&s_full[st]);
    const uint32_t dst = smem_addr_u32(s_rows + st * nvec);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes) : "memory");
    uint32_t off = 0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      if (a.d[b] > 0) {
        const float* src = a.x[b] + row * a.ld[b];
        const uint32_t nbytes = static_cast<uint32_t>(a.d[b]) * 4u;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                     "l"(src), "r"(nbytes), "r"(bar)
                     : "memory");
        off += nbytes;
      }
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int it = 0; it < RING_STAGES - 1; ++it) issue(it);
  }

  int it = 0;
  for (long long row = blockIdx.x; row < a.B; row += gridDim.x, ++it) {
    const int st = it % RING_STAGES;
    ring_mbar_wait(smem_addr_u32(&s_full[st]), static_cast<uint32_t>(it / RING_STAGES) & 1u);
    float4 v[NV];
    float mn = INFINITY, mx = -INFINITY, probe = 0.f;
    const float4* srow = s_rows + st * nvec;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int j = tid + FWD_THREADS * k;
      if (j < nvec) {
        v[k] = srow[j];
        mn = fminf(fminf(mn, v[k].x), fminf(v[k].y, fminf(v[k].z, v[k].w)));
        mx = fmaxf(fmaxf(mx, v[k].x), fmaxf(v[k].y, fmaxf(v[k].z, v[k].w)));
        probe += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      }
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    probe = warp_sum(probe);
    float(*part)[3] = s_part[it & 1];
    if (lane == 0) {
      part[warp][0] = mn;
      part[warp][1] = mx;
      part[warp][2] = probe;
    }
    __syncthreads();  // every thread holds its part of row `it` in registers: the previous row's buffer is free
    if (tid == 0) issue(it + RING_STAGES - 1);
#pragma unroll
    for (int w = 0; w < FWD_THREADS / 32; ++w) {
      mn = fminf(mn, part[w][0]);
      mx = fmaxf(mx, part[w][1]);
    }
    probe = (part[0][2] + part[1][2]) + (part[2][2] + part[3][2]);
    if (probe != probe) mn = mx = __int_as_float(0x7fc00000);
    const float range = __fsub_rn(mx, mn);
    const float inv_range = __frcp_rn(range);
    if (tid == 0) {
      if (a.row_min) a.row_min[row] = mn;
      if (a.row_max) a.row_max[row] = mx;
    }
    const unsigned int grow = static_cast<unsigned int>(a.row0 + static_cast<unsigned long long>(row));
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int j = tid + FWD_THREADS * k;
      if (j < nvec) {
        float f[4] = {__fsub_rn(v[k].x, mn), __fsub_rn(v[k].y, mn), __fsub_rn(v[k].z, mn), __fsub_rn(v[k].w, mn)};
        if (NOISE == PGF_NOISE_PHILOX) {
#pragma unroll
          for (int e = 0; e < 4; ++e) f[e] = __fmul_rn(f[e], inv_range);
          const float4 e4 = s_eps[j];
          const float eh[4] = {e4.x, e4.y, e4.z, e4.w};  // = -ln2 * eps_hat
          const uint4 r = CONSTKEYS ? philox4x32_10_rk(static_cast<unsigned int>(j), grow, PGF_STREAM_LAPLACE, a.offset, a.rk)
                                    : philox4x32_10(static_cast<unsigned int>(j), grow, PGF_STREAM_LAPLACE, a.offset, k0, k1);
          const unsigned int rb[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) f[e] = perturb_fma(f[e], rb[e], eh[e]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) f[e] = __fdiv_rn(f[e], range);
        }
        store_out4<OutT>(a.out, row * a.ld_out + (j << 2), make_float4(f[0], f[1], f[2], f[3]));
      }
    }
  }
}

// Shared-batch variant: every model of an eps x seed sweep perturbs the SAME feature rows (the reference sweep
// re-reads one dataset per run), so a row is fetched, min/max-reduced and normalised once and then perturbed
// once per model with that model's eps_hat row (L1-resident) and Philox key: per model-row the kernel reads
// 4*D/n_models and writes 2*D (bf16) bytes, and the load / reduction / normalisation instructions are amortised.
template <int NV, typename OutT, int RING_STAGES>
__global__ void __launch_bounds__(FWD_THREADS, 5) perturb_fwd_ring_shared_kernel(const PerturbFwdArgs a) {
  extern __shared__ float4 smem4[];
  __shared__ float s_part[2][FWD_THREADS / 32][3];
  __shared__ __align__(8) unsigned long long s_full[RING_STAGES];
  const int nvec = a.D >> 2;
  float4* s_rows = smem4;  // [RING_STAGES][D/4]
  // Philox round keys of every model (key schedule k + r*W), computed once per CTA: 20 words per model, read back as
  // five 128-bit shared loads per model-row instead of 100 integer adds per thread (the schedule of every call)
  uint4* s_rk = reinterpret_cast<uint4*>(smem4 + RING_STAGES * nvec);  // [n_models][5]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < RING_STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&s_full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int m = tid; m < a.n_models; m += FWD_THREADS) {
    const unsigned long long seed = a.model_seeds ? a.model_seeds[m] : a.seed + static_cast<unsigned long long>(m) * a.seed_step;
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_rk + 5 * m);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      dst[2 * r] = k0;
      dst[2 * r + 1] = k1;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
  }
  __syncthreads();
  const uint32_t row_bytes = static_cast<uint32_t>(a.D) * 4u;
  auto issue = [&](int it) {
    const long long row = static_cast<long long>(blockIdx.x) + static_cast<long long>(it) * gridDim.x;
    if (row >= a.B) return;
    const int st = it % RING_STAGES;
    const uint32_t bar = smem_addr_u32(&s_full[st]);
    const uint32_t dst = smem_addr_u32(s_rows + st * nvec);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes) : "memory");
    uint32_t off = 0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      if (a.d[b] > 0) {
        const float* src = a.x[b] + row * a.ld[b];
        const uint32_t nbytes = static_cast<uint32_t>(a.d[b]) * 4u;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                     "l"(src), "r"(nbytes), "r"(bar)
                     : "memory");
        off += nbytes;
      }
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int it = 0; it < RING_STAGES - 1; ++it) issue(it);
  }
  int it = 0;
  for (long long row = blockIdx.x; row < a.B; row += gridDim.x, ++it) {
    const int st = it % RING_STAGES;
    ring_mbar_wait(smem_addr_u32(&s_full[st]), static_cast<uint32_t>(it / RING_STAGES) & 1u);
    float4 v[NV];
    float mn = INFINITY, mx = -INFINITY, probe = 0.f;
    const float4* srow = s_rows + st * nvec;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int j = tid + FWD_THREADS * k;
      if (j < nvec) {
        v[k] = srow[j];
        mn = fminf(fminf(mn, v[k].x), fminf(v[k].y, fminf(v[k].z, v[k].w)));
        mx = fmaxf(fmaxf(mx, v[k].x), fmaxf(v[k].y, fmaxf(v[k].z, v[k].w)));
        probe += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      }
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    probe = warp_sum(probe);
    float(*part)[3] = s_part[it & 1];
    if (lane == 0) {
      part[warp][0] = mn;
      part[warp][1] = mx;
      part[warp][2] = probe;
    }
    __syncthreads();
    if (tid == 0) issue(it + RING_STAGES - 1);
#pragma unroll
    for (int w = 0; w < FWD_THREADS / 32; ++w) {
      mn = fminf(mn, part[w][0]);
      mx = fmaxf(mx, part[w][1]);
    }
    probe = (part[0][2] + part[1][2]) + (part[2][2] + part[3][2]);
    if (probe != probe) mn = mx = __int_as_float(0x7fc00000);
    const float inv_range = __frcp_rn(__fsub_rn(mx, mn));
#pragma unroll
    for (int k = 0; k < NV; ++k) {  // normalised once, perturbed once per model
      v[k].x = __fmul_rn(__fsub_rn(v[k].x, mn), inv_range);
      v[k].y = __fmul_rn(__fsub_rn(v[k].y, mn), inv_range);
      v[k].z = __fmul_rn(__fsub_rn(v[k].z, mn), inv_range);
      v[k].w = __fmul_rn(__fsub_rn(v[k].w, mn), inv_range);
    }
    const unsigned int grow = static_cast<unsigned int>(a.row0 + static_cast<unsigned long long>(row));
    // Models that share a seed draw the SAME noise by definition (same Philox key and counters) -- an eps sweep at one
    // seed, which is how the sweep grid lands on a GPU: the log2-uniforms with the Laplace sign folded in are computed
    // once per run of equal seeds and every model of the run only scales them by its own eps_hat row (one FMA per
    // element).  (lg2 ^ sign) * c == lg2 * (c ^ sign) bit for bit, so the result equals the single-model kernel's.
    float4 sl[NV];
    unsigned int have_lo = 0u, have_hi = 0u;
    bool have = false;
#pragma unroll 1
    for (int m = 0; m < a.n_models; ++m) {
      const uint4 kq0 = s_rk[5 * m];
      if (!have || kq0.x != have_lo || kq0.y != have_hi) {   // first round key pair == the seed
        PhiloxKeys rk;
        rk.k[0] = kq0.x; rk.k[1] = kq0.y; rk.k[2] = kq0.z; rk.k[3] = kq0.w;
#pragma unroll
        for (int q = 1; q < 5; ++q) {
          const uint4 kq = s_rk[5 * m + q];
          rk.k[4 * q] = kq.x; rk.k[4 * q + 1] = kq.y; rk.k[4 * q + 2] = kq.z; rk.k[4 * q + 3] = kq.w;
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int j = tid + FWD_THREADS * k;
          if (j < nvec) {
            const uint4 r = philox4x32_10_rk(static_cast<unsigned int>(j), grow, PGF_STREAM_LAPLACE, a.offset, rk);
            sl[k] = make_float4(signed_lg2_from_bits(r.x), signed_lg2_from_bits(r.y), signed_lg2_from_bits(r.z),
                                signed_lg2_from_bits(r.w));
          }
        }
        have = true;
        have_lo = kq0.x;
        have_hi = kq0.y;
      }
      const float4* ge = reinterpret_cast<const float4*>(a.eps_hat + m * a.s_coef);
      char* outp = static_cast<char*>(a.out) + m * a.s_out * static_cast<long long>(sizeof(OutT));
      if (tid == 0) {
        if (a.row_min) a.row_min[static_cast<long long>(m) * a.B + row] = mn;
        if (a.row_max) a.row_max[static_cast<long long>(m) * a.B + row] = mx;
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int j = tid + FWD_THREADS * k;
        if (j < nvec) {
          const float4 e4 = __ldg(ge + j);
          float4 o;
          o.x = __fmaf_rn(sl[k].x, __fmul_rn(e4.x, -0.69314718055994531f), v[k].x);
          o.y = __fmaf_rn(sl[k].y, __fmul_rn(e4.y, -0.69314718055994531f), v[k].y);
          o.z = __fmaf_rn(sl[k].z, __fmul_rn(e4.z, -0.69314718055994531f), v[k].z);
          o.w = __fmaf_rn(sl[k].w, __fmul_rn(e4.w, -0.69314718055994531f), v[k].w);
          store_out4<OutT>(outp, row * a.ld_out + (j << 2), o);
        }
      }
    }
  }
}

template <int NV, typename OutT>
static int launch_fwd_ring_shared(const PerturbFwdArgs& a, cudaStream_t stream) {
  constexpr int S = 3;
  const size_t smem = static_cast<size_t>(a.D) * sizeof(float) * S + static_cast<size_t>(a.n_models) * 20 * sizeof(uint32_t);
  auto kern = perturb_fwd_ring_shared_kernel<NV, OutT, S>;
  ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);
  const int occ = cached_occupancy(reinterpret_cast<const void*>(kern), FWD_THREADS, smem, 2);
  int gx = num_sms() * occ;
  if (gx > a.B) gx = a.B;
  kern<<<gx, FWD_THREADS, smem, stream>>>(a);
  PGF_CUDA_LAUNCH_CHECK("pgf_perturb_gate_fwd(shared ring)");
  return PGF_OK;
}

template <int NV, int NOISE, typename OutT, int RING_STAGES>
static int launch_fwd_ring_s(const PerturbFwdArgs& a, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(a.D) * sizeof(float) * (RING_STAGES + (NOISE == PGF_NOISE_NONE ? 0 : 1));
  auto kern = perturb_fwd_ring_kernel<NV, NOISE, OutT, true, RING_STAGES>;
  ensure_dynamic_smem(reinterpret_cast<const void*>(kern), smem);
  const int occ = cached_occupancy(reinterpret_cast<const void*>(kern), FWD_THREADS, smem, 2);
  int gx = num_sms() * occ;
  if (gx > a.B) gx = a.B;
  kern<<<gx, FWD_THREADS, smem, stream>>>(a);
  PGF_CUDA_LAUNCH_CHECK("pgf_perturb_gate_fwd(ring)");
  return PGF_OK;
}

template <int NV, int NOISE, typename OutT>
static int launch_fwd_ring(const PerturbFwdArgs& a, cudaStream_t stream) {
  static const char* env = getenv("PGF_RING_STAGES");  // tuning knob: rows in flight per CTA = stages - 1
  if (env && env[0] == '4') return launch_fwd_ring_s<NV, NOISE, OutT, 4>(a, stream);
  return launch_fwd_ring_s<NV, NOISE, OutT, 3>(a, stream);
}

template <int NV>
static int launch_fwd_ring_nv(const PerturbFwdArgs& a, int noise, int out_dtype, cudaStream_t s) {
  if (noise == PGF_NOISE_PHILOX)
    return out_dtype == PGF_DT_F32 ? launch_fwd_ring<NV, PGF_NOISE_PHILOX, float>(a, s)
                                   : launch_fwd_ring<NV, PGF_NOISE_PHILOX, __nv_bfloat16>(a, s);
  return out_dtype == PGF_DT_F32 ? launch_fwd_ring<NV, PGF_NOISE_NONE, float>(a, s)
                                 : launch_fwd_ring<NV, PGF_NOISE_NONE, __nv_bfloat16>(a, s);
}

template <typename K>
static int persistent_grid(K kernel, int threads, size_t smem, int n_models, int B) {
  // persistent CTAs: exactly one resident wave (SMs x occupancy), each CTA looping over rows
  const int occ = cached_occupancy(reinterpret_cast<const void*>(kernel), threads, smem, 4);
  int gx = (num_sms() * occ) / (n_models > 0 ? n_models : 1);
  if (gx > B) gx = B;
  if (gx < 1) gx = 1;
  return gx;
}

template <int NV, int NOISE, typename OutT>
static int launch_fwd_gate(const PerturbFwdArgs& a, bool want_gate, cudaStream_t stream) {
  const size_t smem = (NOISE == PGF_NOISE_NONE) ? 0 : static_cast<size_t>(a.D) * sizeof(float) * (want_gate ? 2 : 1);
  const bool constkeys = NOISE == PGF_NOISE_PHILOX && a.n_models == 1 && !want_gate;
#define PGF_FWD_LAUNCH(GATE, CK)                                                                                  \
  do {                                                                                                            \
    auto kern = perturb_gate_fwd_kernel<NV, NOISE, OutT, GATE, CK>;                                               \
    const dim3 grid(persistent_grid(kern, FWD_THREADS, smem, a.n_models, a.B * (a.n_rep > 1 ? a.n_rep : 1)), a.n_models); \
    launch(kern, grid, dim3(FWD_THREADS), smem, stream, a);                                                       \
  } while (0)
  if (want_gate) PGF_FWD_LAUNCH(true, false);
  else if (constkeys) PGF_FWD_LAUNCH(false, (NOISE == PGF_NOISE_PHILOX));
  else PGF_FWD_LAUNCH(false, false);
#undef PGF_FWD_LAUNCH
  PGF_CUDA_LAUNCH_CHECK("pgf_perturb_gate_fwd");
  return PGF_OK;
}

template <int NV>
static int launch_fwd_nv(const PerturbFwdArgs& a, int noise, int out_dtype, bool want_gate, cudaStream_t s) {
#define PGF_DISPATCH_OUT(NOISE)                                                              \
  return out_dtype == PGF_DT_F32 ? launch_fwd_gate<NV, NOISE, float>(a, want_gate, s)         \
                                 : launch_fwd_gate<NV, NOISE, __nv_bfloat16>(a, want_gate, s)
  switch (noise) {
    case PGF_NOISE_INJECTED: PGF_DISPATCH_OUT(PGF_NOISE_INJECTED);
    case PGF_NOISE_PHILOX: PGF_DISPATCH_OUT(PGF_NOISE_PHILOX);
    default: PGF_DISPATCH_OUT(PGF_NOISE_NONE);
  }
#undef PGF_DISPATCH_OUT
}

static int perturb_gate_fwd_one(const PerturbFwdArgs& a, int noise, int out_dtype, bool want_gate, cudaStream_t s) {
  const int nv = (a.D / 4 + FWD_THREADS - 1) / FWD_THREADS;
  static const bool no_ring = getenv("PGF_PERTURB_NO_RING") != nullptr;
  if (!no_ring && noise != PGF_NOISE_INJECTED && !want_gate && a.n_models == 1 && !a.st && !a.gather && a.n_rep <= 1 &&
      static_cast<long long>(a.B) * a.D >= (1LL << 22)) {
    if (nv <= 2) return launch_fwd_ring_nv<2>(a, noise, out_dtype, s);
    if (nv <= 5) return launch_fwd_ring_nv<5>(a, noise, out_dtype, s);
    if (nv <= 8) return launch_fwd_ring_nv<8>(a, noise, out_dtype, s);
  }
  if (nv <= 2) return launch_fwd_nv<2>(a, noise, out_dtype, want_gate, s);
  if (nv <= 5) return launch_fwd_nv<5>(a, noise, out_dtype, want_gate, s);
  if (nv <= 8) return launch_fwd_nv<8>(a, noise, out_dtype, want_gate, s);
  set_error("pgf_perturb_gate_fwd: fused width D=%d exceeds the register-resident limit 4096", a.D);
  return PGF_ERR_UNSUPPORTED;
}

int perturb_gate_fwd(const PerturbFwdArgs& a_in, int noise, int out_dtype, bool want_gate, cudaStream_t s) {
  PerturbFwdArgs a = a_in;
  a.rk = philox_make_keys(a.seed);
  // Large batches: one launch per model, so the Philox round keys are compile-time-indexed kernel
  // arguments (constant-bank operands).  Small batches (the B=8 sweep): one grouped launch.
  const bool big = static_cast<long long>(a.B) * a.D >= (1LL << 22);
  static const bool no_ring = getenv("PGF_PERTURB_NO_RING") != nullptr;
  if (a.st || a.gather || a.n_rep > 1) return perturb_gate_fwd_one(a, noise, out_dtype, want_gate, s);
  if (!no_ring && noise == PGF_NOISE_PHILOX && a.n_models > 1 && !want_gate && big && a.sx[0] == 0 && a.sx[1] == 0 && a.sx[2] == 0) {
    // one batch shared by the whole sweep: fetch / normalise each row once, perturb it once per model
    const int nv = (a.D / 4 + FWD_THREADS - 1) / FWD_THREADS;
    const bool f32 = out_dtype == PGF_DT_F32;
    if (nv <= 2) return f32 ? launch_fwd_ring_shared<2, float>(a, s) : launch_fwd_ring_shared<2, __nv_bfloat16>(a, s);
    if (nv <= 5) return f32 ? launch_fwd_ring_shared<5, float>(a, s) : launch_fwd_ring_shared<5, __nv_bfloat16>(a, s);
    if (nv <= 8) return f32 ? launch_fwd_ring_shared<8, float>(a, s) : launch_fwd_ring_shared<8, __nv_bfloat16>(a, s);
  }
  const bool split = noise == PGF_NOISE_PHILOX && a.n_models > 1 && !want_gate && !a.model_seeds && big;
  if (!split) return perturb_gate_fwd_one(a, noise, out_dtype, want_gate, s);
  const size_t esz = out_dtype == PGF_DT_F32 ? 4 : 2;
  for (int m = 0; m < a_in.n_models; ++m) {
    PerturbFwdArgs b = a_in;
    b.n_models = 1;
    for (int i = 0; i < 3; ++i)
      if (b.x[i]) b.x[i] = a_in.x[i] + m * a_in.sx[i];
    b.w = a_in.w + m * a_in.s_coef;
    b.eps_hat = a_in.eps_hat + m * a_in.s_coef;
    b.out = static_cast<char*>(a_in.out) + static_cast<size_t>(m) * a_in.s_out * esz;
    if (b.row_min) b.row_min = a_in.row_min + static_cast<long long>(m) * a_in.B;
    if (b.row_max) b.row_max = a_in.row_max + static_cast<long long>(m) * a_in.B;
    b.seed = a_in.seed + static_cast<unsigned long long>(m) * a_in.seed_step;
    b.rk = philox_make_keys(b.seed);
    const int rc = perturb_gate_fwd_one(b, noise, out_dtype, want_gate, s);
    if (rc != PGF_OK) return rc;
  }
  return PGF_OK;
}

// ------------------------------------------------------------------------------------------
// backward wrt DP:  dDP[d] = deps_dDP[d] * sum_b dF[b,d] * lap[b,d]
// (reference: autograd through models.py:75-76; the gate's own contribution is zero in exact
//  arithmetic because both mask planes multiply the same feature -- SURVEY.md section 0 item 4)
// Stage 1: each CTA column-reduces a slab of rows for 512 columns into partial[slab][D].
// Stage 2: deterministic sum over slabs, times the per-column coefficient.
// ------------------------------------------------------------------------------------------
struct PerturbBwdArgs {
  const void* dF;
  long long ld;
  long long s_dF;
  int B, D;
  const float* lap;
  unsigned long long seed;
  unsigned long long seed_step;
  const unsigned long long* model_seeds;
  int nslab;
  PhiloxKeys rk;
  unsigned int offset;
  unsigned long long row0;
  int rows_per_slab;
  float* partial;  // [nslab, D]
  // nslab == 1 (B <= 32, the reference's batch): the kernel applies the finalize step itself, no second launch
  const float* coef; long long s_coef;
  float* dDP; long long s_dDP;
  int direct, accumulate;
  const StepState* st;      // offset += st->noise_offset; Adam coefficients of the DP group
  int fused;                // direct mode only: Adam(DP) + coefficient refresh in the same thread (DpAdamFuse)
  DpAdamFuse f;
};

// (w, eps_hat, d eps_hat/d DP) of one column: the arithmetic of dp_coeffs_kernel, shared with the fused DP-pass tail
__device__ __forceinline__ void dp_coeff_one(float x, float exp_eps, int fixed, float& w, float& eh, float& de) {
  w = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
  const float num = __fsub_rn(exp_eps, w);
  const float ratio = __fdiv_rn(num, __fsub_rn(1.0f, w));
  const float L = logf(ratio);
  eh = fixed ? __fdiv_rn(1.0f, L) : L;
  const float t = __fdiv_rn(__fmul_rn(__fsub_rn(exp_eps, 1.0f), w), num);
  de = fixed ? __fdiv_rn(-t, __fmul_rn(L, L)) : t;
}

template <typename InT>
__device__ __forceinline__ float4 load_in4(const void* p, long long off);
template <>
__device__ __forceinline__ float4 load_in4<float>(const void* p, long long off) {
  return ldg_stream(reinterpret_cast<const float4*>(static_cast<const float*>(p) + off));
}
template <>
__device__ __forceinline__ float4 load_in4<__nv_bfloat16>(const void* p, long long off) {
  const uint2 u = ldg_stream_u2(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p) + off));
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

template <int NOISE, typename InT, bool CONSTKEYS>
__global__ void __launch_bounds__(128) perturb_bwd_dp_kernel(const PerturbBwdArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;  // float4 column index
  const int col = j << 2;
  griddep_wait();
  griddep_launch();
  if (col >= a.D) return;
  const unsigned int offs = a.offset + (a.st ? static_cast<unsigned int>(a.st->noise_offset) : 0u);
  const int slab = blockIdx.y, model = blockIdx.z;
  const int r0 = slab * a.rows_per_slab;
  const int r1 = min(a.B, r0 + a.rows_per_slab);
  const unsigned long long seed = a.model_seeds ? a.model_seeds[model] : a.seed + static_cast<unsigned long long>(model) * a.seed_step;
  const unsigned int k0 = static_cast<unsigned int>(seed), k1 = static_cast<unsigned int>(seed >> 32);
  const void* dFm = static_cast<const char*>(a.dF) + model * a.s_dF * static_cast<long long>(sizeof(InT));
  const float* lapm = a.lap ? a.lap + static_cast<long long>(model) * a.B * a.D : nullptr;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int U = 8;
  int r = r0;
  for (; r + U <= r1; r += U) {
    float4 g[U], l[U];
#pragma unroll
    for (int u = 0; u < U; ++u) g[u] = load_in4<InT>(dFm, static_cast<long long>(r + u) * a.ld + col);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (NOISE == PGF_NOISE_INJECTED) {
        l[u] = ldg_stream(reinterpret_cast<const float4*>(lapm + static_cast<long long>(r + u) * a.D + col));
      } else {
        const uint4 q = CONSTKEYS ? philox4x32_10_rk(static_cast<unsigned int>(j), static_cast<unsigned int>(a.row0 + r + u),
                                                     PGF_STREAM_LAPLACE, offs, a.rk)
                                 : philox4x32_10(static_cast<unsigned int>(j), static_cast<unsigned int>(a.row0 + r + u),
                                                 PGF_STREAM_LAPLACE, offs, k0, k1);
        l[u] = make_float4(laplace_from_bits(q.x), laplace_from_bits(q.y), laplace_from_bits(q.z),
                           laplace_from_bits(q.w));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      acc.x = fmaf(g[u].x, l[u].x, acc.x);
      acc.y = fmaf(g[u].y, l[u].y, acc.y);
      acc.z = fmaf(g[u].z, l[u].z, acc.z);
      acc.w = fmaf(g[u].w, l[u].w, acc.w);
    }
  }
  for (; r < r1; ++r) {
    const float4 g = load_in4<InT>(dFm, static_cast<long long>(r) * a.ld + col);
    float4 l;
    if (NOISE == PGF_NOISE_INJECTED) {
      l = ldg_stream(reinterpret_cast<const float4*>(lapm + static_cast<long long>(r) * a.D + col));
    } else {
      const uint4 q = CONSTKEYS ? philox4x32_10_rk(static_cast<unsigned int>(j), static_cast<unsigned int>(a.row0 + r),
                                                   PGF_STREAM_LAPLACE, offs, a.rk)
                               : philox4x32_10(static_cast<unsigned int>(j), static_cast<unsigned int>(a.row0 + r),
                                               PGF_STREAM_LAPLACE, offs, k0, k1);
      l = make_float4(laplace_from_bits(q.x), laplace_from_bits(q.y), laplace_from_bits(q.z), laplace_from_bits(q.w));
    }
    acc.x = fmaf(g.x, l.x, acc.x);
    acc.y = fmaf(g.y, l.y, acc.y);
    acc.z = fmaf(g.z, l.z, acc.z);
    acc.w = fmaf(g.w, l.w, acc.w);
  }
  if (a.direct) {   // what perturb_bwd_dp_finalize_kernel computes from a single slab, bit for bit
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
    float* o = a.dDP + model * a.s_dDP + col;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (a.coef) v[q] = __fmul_rn(v[q], a.coef[model * a.s_coef + col + q]);
      v[q] = a.accumulate ? __fadd_rn(o[q], v[q]) : v[q];
      o[q] = v[q];
    }
    if (a.fused) {   // DP_optimizer.step() (past_acc.py:203) and the coefficient rows the next forward reads
      const AdamCoef c = adam_coef_at(a.f.c, a.st, 0);
      const long long i = static_cast<long long>(model) * a.D + col;
      const float ee = a.f.exp_eps[model];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float p = a.f.DP[i + q], m = a.f.DP_m[i + q], vv = a.f.DP_v[i + q];
        adam_update(p, m, vv, v[q], c);
        a.f.DP[i + q] = p; a.f.DP_m[i + q] = m; a.f.DP_v[i + q] = vv;
        float w, eh, de;
        dp_coeff_one(p, ee, a.f.fixed, w, eh, de);
        const long long ci = model * a.s_coef + col + q;
        a.f.w[ci] = w; a.f.eps_hat[ci] = eh; a.f.deps[ci] = de;
      }
    }
    return;
  }
  *reinterpret_cast<float4*>(a.partial + (static_cast<long long>(model) * a.nslab + slab) * a.D + col) = acc;
}

// Deterministic column reduction of a [nslab, D] partial buffer: a CTA owns 32 columns, its warps take
// interleaved slabs (4 loads in flight each), then combine in a fixed order.
constexpr int FIN_WARPS = 32;
__global__ void __launch_bounds__(FIN_WARPS * 32) perturb_bwd_dp_finalize_kernel(const float* __restrict__ partial, int nslab, int D,
                                                                                const float* __restrict__ coef, long long s_coef,
                                                                                float* __restrict__ dDP, long long s_dDP,
                                                                                float accumulate) {
  __shared__ float s_part[FIN_WARPS][32];
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + lane;
  const int model = blockIdx.y;
  const float* P = partial + static_cast<long long>(model) * nslab * D + d;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (d < D) {
    int i = rg;
    for (; i + 3 * FIN_WARPS < nslab; i += 4 * FIN_WARPS) {
      s0 += P[static_cast<long long>(i) * D];
      s1 += P[static_cast<long long>(i + FIN_WARPS) * D];
      s2 += P[static_cast<long long>(i + 2 * FIN_WARPS) * D];
      s3 += P[static_cast<long long>(i + 3 * FIN_WARPS) * D];
    }
    for (; i < nslab; i += FIN_WARPS) s0 += P[static_cast<long long>(i) * D];
  }
  s_part[rg][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (rg == 0 && d < D) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < FIN_WARPS; ++w) s += s_part[w][lane];
    const float v = coef ? __fmul_rn(s, coef[model * s_coef + d]) : s;   // explicit roundings: the single-slab path of
    float* o = dDP + model * s_dDP + d;                                   // perturb_bwd_dp_kernel must match bit for bit
    *o = accumulate != 0.f ? __fadd_rn(*o, v) : v;
  }
}

// out[n] = (coef ? coef[n] : 1) * sum_r partial[r][n], fixed summation order (the column partials of the GEMM epilogues)
int reduce_partials(const float* partial, int rows, int N, const float* coef, float* out, int accumulate, cudaStream_t s) {
  perturb_bwd_dp_finalize_kernel<<<dim3((N + 31) / 32, 1), FIN_WARPS * 32, 0, s>>>(partial, rows, N, coef, 0, out, 0, accumulate ? 1.f : 0.f);
  PGF_CUDA_LAUNCH_CHECK("pgf_reduce_partials");
  return PGF_OK;
}

int perturb_bwd_slabs(int B, int D, int n_models) {
  const int col_ctas = (D / 4 + 127) / 128 * (n_models > 0 ? n_models : 1);
  int slabs = (num_sms() * 8 + col_ctas - 1) / col_ctas;  // ~8 CTAs of 128 threads per SM
  const int max_slabs = (B + 31) / 32;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  return slabs;
}

int perturb_gate_bwd_dp(const void* dF, int dtype, long long ld, long long s_dF, int B, int D, int n_models, int noise,
                        const float* lap, unsigned long long seed, unsigned long long seed_step,
                        const unsigned long long* model_seeds, unsigned int offset, unsigned long long row0, const float* coef,
                        long long s_coef, float* workspace, size_t workspace_bytes, float* dDP, long long s_dDP, int accumulate,
                        cudaStream_t s, const StepState* st, const DpAdamFuse* fuse) {
  const int slabs = perturb_bwd_slabs(B, D, n_models);
  if (fuse && slabs != 1) {
    set_error("pgf_perturb_gate_bwd_dp: the fused Adam(DP) tail needs a single-slab launch (B <= 32), got B=%d", B);
    return PGF_ERR_UNSUPPORTED;
  }
  const size_t need = static_cast<size_t>(n_models) * slabs * D * sizeof(float);
  if (workspace_bytes < need) {
    set_error("pgf_perturb_gate_bwd_dp: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return PGF_ERR_WORKSPACE;
  }
  PerturbBwdArgs a;
  a.dF = dF;
  a.ld = ld;
  a.s_dF = s_dF;
  a.B = B;
  a.D = D;
  a.lap = lap;
  a.seed = seed;
  a.seed_step = seed_step;
  a.model_seeds = model_seeds;
  a.nslab = slabs;
  a.offset = offset;
  a.row0 = row0;
  a.rows_per_slab = (B + slabs - 1) / slabs;
  a.partial = workspace;
  a.direct = slabs == 1;
  a.coef = coef; a.s_coef = s_coef; a.dDP = dDP; a.s_dDP = s_dDP; a.accumulate = accumulate;
  a.st = st; a.fused = fuse != nullptr;
  if (fuse) a.f = *fuse; else a.f = DpAdamFuse{};
  const dim3 grid((D / 4 + 127) / 128, slabs, n_models);
  a.rk = philox_make_keys(seed);
  if (noise == PGF_NOISE_INJECTED) {
    if (dtype == PGF_DT_F32)
      launch(perturb_bwd_dp_kernel<PGF_NOISE_INJECTED, float, false>, grid, dim3(128), 0, s, a);
    else
      launch(perturb_bwd_dp_kernel<PGF_NOISE_INJECTED, __nv_bfloat16, false>, grid, dim3(128), 0, s, a);
  } else if (n_models == 1 && !model_seeds) {
    if (dtype == PGF_DT_F32)
      launch(perturb_bwd_dp_kernel<PGF_NOISE_PHILOX, float, true>, grid, dim3(128), 0, s, a);
    else
      launch(perturb_bwd_dp_kernel<PGF_NOISE_PHILOX, __nv_bfloat16, true>, grid, dim3(128), 0, s, a);
  } else {
    if (dtype == PGF_DT_F32)
      launch(perturb_bwd_dp_kernel<PGF_NOISE_PHILOX, float, false>, grid, dim3(128), 0, s, a);
    else
      launch(perturb_bwd_dp_kernel<PGF_NOISE_PHILOX, __nv_bfloat16, false>, grid, dim3(128), 0, s, a);
  }
  PGF_CUDA_LAUNCH_CHECK("pgf_perturb_gate_bwd_dp");
  if (a.direct) return PGF_OK;
  const dim3 fgrid((D + 31) / 32, n_models);
  perturb_bwd_dp_finalize_kernel<<<fgrid, FIN_WARPS * 32, 0, s>>>(workspace, slabs, D, coef, s_coef, dDP, s_dDP, accumulate ? 1.f : 0.f);
  PGF_CUDA_LAUNCH_CHECK("pgf_perturb_gate_bwd_dp(finalize)");
  return PGF_OK;
}

// ------------------------------------------------------------------------------------------
// per-column coefficients from DP (models.py:73,75): w = sigmoid(DP),
//   eps_hat = 1/log((e^eps - w)/(1 - w))    [fixed]      or   log(...)   [unfixed, model.py:57]
//   deps_dDP = d eps_hat / d DP = d eps_hat/dw * w(1-w)
//     fixed:   -(E-1) w / (L^2 (E-w)),    unfixed:  (E-1) w / (E-w),   L = log((E-w)/(1-w))
// ------------------------------------------------------------------------------------------
__global__ void dp_coeffs_kernel(const float* __restrict__ DP, const float* __restrict__ exp_eps_arr, int fixed, int D,
                                 float* __restrict__ w_out, float* __restrict__ eps_hat, float* __restrict__ deps) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int model = blockIdx.y;
  if (d >= D) return;
  const long long i = static_cast<long long>(model) * D + d;
  float w, eh, de;
  dp_coeff_one(DP[i], exp_eps_arr[model], fixed, w, eh, de);
  if (w_out) w_out[i] = w;
  if (eps_hat) eps_hat[i] = eh;
  if (deps) deps[i] = de;
}

int dp_coeffs(const float* DP, const float* exp_eps, int fixed, int D, int n_models, float* w, float* eps_hat, float* deps,
              cudaStream_t s) {
  const dim3 grid((D + 255) / 256, n_models);
  dp_coeffs_kernel<<<grid, 256, 0, s>>>(DP, exp_eps, fixed, D, w, eps_hat, deps);
  PGF_CUDA_LAUNCH_CHECK("pgf_dp_coeffs");
  return PGF_OK;
}

// ------------------------------------------------------------------------------------------
// backward of the min-max normalisation wrt the raw feature blocks (only needed when the
// encoders above the head are trained, reference: autograd through models.py:70-72):
//   n = (x-mn)/r, r = mx-mn;  dx_j = dn_j/r - [j==argmin] sum_k dn_k (1-n_k)/r - [j==argmax] sum_k dn_k n_k / r
// torch.min/max(dim) route the gradient to the single index they return (first occurrence).
// ------------------------------------------------------------------------------------------

template <int NV, typename InT>
__global__ void __launch_bounds__(256) minmax_norm_bwd_kernel(const NormBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const int d01 = a.d[0] + a.d[1];
  for (long long row = static_cast<long long>(blockIdx.x) * warps_per_cta + (threadIdx.x >> 5); row < a.B;
       row += static_cast<long long>(gridDim.x) * warps_per_cta) {
    float4 v[NV];
    float mn = INFINITY, mx = -INFINITY;
    int imn = 0x7fffffff, imx = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int col = (lane + 32 * k) << 2;
      if (col < a.D) {
        const float* p = col < a.d[0] ? a.x[0] + row * a.ld[0] + col
                                      : (col < d01 ? a.x[1] + row * a.ld[1] + (col - a.d[0])
                                                   : a.x[2] + row * a.ld[2] + (col - d01));
        v[k] = *reinterpret_cast<const float4*>(p);
        const float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (e[q] < mn) { mn = e[q]; imn = col + q; }
          if (e[q] > mx) { mx = e[q]; imx = col + q; }
        }
      }
    }
    // warp arg-min / arg-max, first occurrence on ties
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float omn = __shfl_xor_sync(0xffffffffu, mn, o);
      const int oimn = __shfl_xor_sync(0xffffffffu, imn, o);
      if (omn < mn || (omn == mn && oimn < imn)) { mn = omn; imn = oimn; }
      const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oimx = __shfl_xor_sync(0xffffffffu, imx, o);
      if (omx > mx || (omx == mx && oimx < imx)) { mx = omx; imx = oimx; }
    }
    const float r = mx - mn;
    const float inv_r = 1.0f / r;
    float4 g[NV];
    float s_min = 0.f, s_max = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int col = (lane + 32 * k) << 2;
      if (col < a.D) {
        g[k] = load_in4<InT>(a.dn, row * a.ld_dn + col);
        const float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
        const float ge[4] = {g[k].x, g[k].y, g[k].z, g[k].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float n = (e[q] - mn) * inv_r;
          s_max += ge[q] * n;
          s_min += ge[q] * (1.0f - n);
        }
      }
    }
    s_min = warp_sum(s_min) * inv_r;
    s_max = warp_sum(s_max) * inv_r;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int col = (lane + 32 * k) << 2;
      if (col < a.D) {
        float o[4] = {g[k].x * inv_r, g[k].y * inv_r, g[k].z * inv_r, g[k].w * inv_r};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (col + q == imn) o[q] -= s_min;
          if (col + q == imx) o[q] -= s_max;
        }
        float* p = col < a.d[0] ? a.dx[0] + row * a.ld_dx[0] + col
                                : (col < d01 ? a.dx[1] + row * a.ld_dx[1] + (col - a.d[0])
                                             : a.dx[2] + row * a.ld_dx[2] + (col - d01));
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

template <int NV>
static int launch_norm_bwd(const NormBwdArgs& a, int dtype, cudaStream_t s) {
  int grid = num_sms() * 2;
  const int max_grid = (a.B + 7) / 8;
  if (grid > max_grid) grid = max_grid;
  if (grid < 1) grid = 1;
  if (dtype == PGF_DT_F32)
    minmax_norm_bwd_kernel<NV, float><<<grid, 256, 0, s>>>(a);
  else
    minmax_norm_bwd_kernel<NV, __nv_bfloat16><<<grid, 256, 0, s>>>(a);
  PGF_CUDA_LAUNCH_CHECK("pgf_minmax_norm_bwd");
  return PGF_OK;
}

int minmax_norm_bwd(const NormBwdArgs& a, int dtype, cudaStream_t s) {
  const int nv = (a.D / 4 + 31) / 32;
  if (nv <= 8) return launch_norm_bwd<8>(a, dtype, s);
  if (nv <= 20) return launch_norm_bwd<20>(a, dtype, s);
  if (nv <= 32) return launch_norm_bwd<32>(a, dtype, s);
  set_error("pgf_minmax_norm_bwd: fused width D=%d exceeds the register-resident limit 4096", a.D);
  return PGF_ERR_UNSUPPORTED;
}

}  // namespace pgf
