// The reference's OLDER "PriGumbel" head tail (SURVEY.md section 8 row a-alt), between fc2 and the classifier:
//   gumbel_dropout  train_val.py:95-101   r = z * gumbel_softmax([w, 1-w], tau, hard)[:,1] / (1 - w)   (w [H] shared by the batch)
//   Lap_noise       train_val.py:114-123  out = (r - min_row) / (max_row - min_row) + Laplace(0, 1/eps)  (ONE scalar per row)
//   loss_function   train_val.py:80-93    the w-dependent term max_j((1 - w_j) e^eps + w_j) and its gradient
// and their autograd.  Unlike the main path the normalised tensor is an ACTIVATION here (fc2 is trained through it), so the
// backward carries the full min-max gradient (arg-min / arg-max routing, first occurrence on ties like torch.min / torch.max).
// The op is HBM-bound and tiny at the reference's batch of 8; rows are warp-resident (H <= 2048), one pass per direction.
#include <type_traits>

#include "pgf_kernels.cuh"
#include "philox.cuh"

namespace pgf {

namespace {

constexpr int PG_THREADS = 256;
constexpr int PG_WARPS = PG_THREADS / 32;

// ---- per-column coefficients: one CTA --------------------------------------------------------------------------------
// coef [4,H]: 0 mask_j (soft y1, or the straight-through composite (onehot1 - y1) + y1), 1 (1 - w_j), 2 d mask_j / d w_j
// (softmax backward of both planes, torch's operation order), 3 1/(1-w) (only used for vanishing numerators, see div_rn_z).  wloss[0] = max_j((1-w_j) e^eps + w_j), wloss[1] = argmax.
__global__ void __launch_bounds__(PG_THREADS) prigumbel_coef_kernel(const float* __restrict__ w, const float* __restrict__ gum,
                                                                    int H, float exp_eps, float tau, int hard,
                                                                    unsigned int k0, unsigned int k1, unsigned int offset,
                                                                    float* __restrict__ coef, float* __restrict__ wloss) {
  __shared__ float s_v[PG_WARPS];
  __shared__ int s_i[PG_WARPS];
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int j = threadIdx.x; j < H; j += PG_THREADS) {
    const float wj = w[j];
    float g0, g1;
    if (gum != nullptr) {
      g0 = gum[2 * j];
      g1 = gum[2 * j + 1];
    } else {
      const uint4 q0 = philox4x32_10(static_cast<unsigned int>(j >> 2), 0u, PGF_STREAM_GUMBEL0, offset, k0, k1);
      const uint4 q1 = philox4x32_10(static_cast<unsigned int>(j >> 2), 0u, PGF_STREAM_GUMBEL1, offset, k0, k1);
      const unsigned int a[4] = {q0.x, q0.y, q0.z, q0.w}, b[4] = {q1.x, q1.y, q1.z, q1.w};
      g0 = gumbel_from_bits(a[j & 3]);
      g1 = gumbel_from_bits(b[j & 3]);
    }
    const float omw = 1.0f - wj;
    const float a0 = __fdiv_rn(wj + g0, tau), a1 = __fdiv_rn(omw + g1, tau);   // F.gumbel_softmax: (logits + gumbels) / tau
    const float m = fmaxf(a0, a1);
    const float e0 = expf(a0 - m), e1 = expf(a1 - m);
    const float sum = e0 + e1;
    const float y0 = __fdiv_rn(e0, sum), y1 = __fdiv_rn(e1, sum);
    const float mask = hard ? ((y1 > y0 ? 1.0f : 0.0f) - y1) + y1 : y1;          // y_hard - y_soft.detach() + y_soft
    // softmax backward with upstream (0, g): d a0 = -y0 y1 g, d a1 = y1 (g - y1 g); logits = (w, 1 - w)
    const float dmask = __fdiv_rn(-(y0 * y1), tau) - __fdiv_rn(y1 - y1 * y1, tau);
    coef[j] = mask;
    coef[H + j] = omw;
    coef[2 * H + j] = dmask;
    coef[3 * H + j] = __frcp_rn(omw);
    const float t = omw * exp_eps + wj;
    if (t > best) { best = t; besti = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < PG_WARPS; ++k)
      if (s_v[k] > best || (s_v[k] == best && s_i[k] < besti)) { best = s_v[k]; besti = s_i[k]; }
    wloss[0] = best;
    wloss[1] = static_cast<float>(besti);
  }
}

// IEEE division with a short-cut for zero / vanishing numerators: at the reference's tau = 0.01 the gate is 0 on about half of the
// columns, and 0 / x takes the slow path of __fdiv_rn (the same effect as in the Adam kernel).
// 0 / x = 0 with the numerator's sign for any finite positive or negative x != 0; x = 0 or NaN falls through to the division.
// In train mode (soft gate) the same columns carry masks like e^-95: the products are denormal, which is the other slow path
// of __fdiv_rn.  Below 2^-100 the quotient is formed as num * (1/den) (FMUL handles denormals at full rate); after the row
// min-max normalisation such an entry is indistinguishable from 0 in fp32 either way.
__device__ __forceinline__ float div_rn_z(float num, float den, float inv_den) {
  return (fabsf(num) < 7.9e-31f && den > 0.f) ? num * inv_den : __fdiv_rn(num, den);
}

__device__ __forceinline__ float div0(float num, float den) { return (num == 0.f && den > 0.f) ? num : __fdiv_rn(num, den); }

// r = (z * mask) / (1 - w), the reference's operation order; row arg-min / arg-max, first occurrence on ties.
template <int NV>
__device__ __forceinline__ void dropout_row(const float* __restrict__ zrow, const float* __restrict__ coef, int H, int lane,
                                            float4 (&z)[NV], float4 (&r)[NV], float& mn, float& mx, int& imn, int& imx) {
  mn = INFINITY; mx = -INFINITY; imn = 0x7fffffff; imx = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int col = (lane + 32 * k) << 2;
    if (col < H) {
      z[k] = *reinterpret_cast<const float4*>(zrow + col);
      const float4 m4 = *reinterpret_cast<const float4*>(coef + col);
      const float4 o4 = *reinterpret_cast<const float4*>(coef + H + col);
      const float4 i4 = *reinterpret_cast<const float4*>(coef + 3 * H + col);
      r[k] = make_float4(div_rn_z(z[k].x * m4.x, o4.x, i4.x), div_rn_z(z[k].y * m4.y, o4.y, i4.y),
                         div_rn_z(z[k].z * m4.z, o4.z, i4.z), div_rn_z(z[k].w * m4.w, o4.w, i4.w));
      const float e[4] = {r[k].x, r[k].y, r[k].z, r[k].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (e[q] < mn) { mn = e[q]; imn = col + q; }
        if (e[q] > mx) { mx = e[q]; imx = col + q; }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float omn = __shfl_xor_sync(0xffffffffu, mn, o);
    const int oimn = __shfl_xor_sync(0xffffffffu, imn, o);
    if (omn < mn || (omn == mn && oimn < imn)) { mn = omn; imn = oimn; }
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oimx = __shfl_xor_sync(0xffffffffu, imx, o);
    if (omx > mx || (omx == mx && oimx < imx)) { mx = omx; imx = oimx; }
  }
}

template <int NV>
__global__ void __launch_bounds__(PG_THREADS) prigumbel_fwd_kernel(const PriGumbelArgs a) {
  const int lane = threadIdx.x & 31;
  for (long long row = static_cast<long long>(blockIdx.x) * PG_WARPS + (threadIdx.x >> 5); row < a.B;
       row += static_cast<long long>(gridDim.x) * PG_WARPS) {
    float4 z[NV], r[NV];
    float mn, mx;
    int imn, imx;
    dropout_row<NV>(a.z + row * a.ldz, a.coef, a.H, lane, z, r, mn, mx, imn, imx);
    float noise;
    if (a.lap != nullptr) {
      noise = a.lap[row];
    } else {
      const uint4 q = philox4x32_10(0u, static_cast<unsigned int>(a.row0 + row), PGF_STREAM_LAPLACE, a.offset, a.k0, a.k1);
      noise = laplace_from_bits(q.x) * a.inv_eps;                                  // Laplace(0, 1/eps)
    }
    const float range = mx - mn;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int col = (lane + 32 * k) << 2;
      if (col < a.H) {
        // (the row's minimum itself is an exact 0 / range: same short-cut)
        const float4 o = make_float4(div0(r[k].x - mn, range) + noise, div0(r[k].y - mn, range) + noise,
                                     div0(r[k].z - mn, range) + noise, div0(r[k].w - mn, range) + noise);
        *reinterpret_cast<float4*>(a.out + row * a.ld_out + col) = o;
      }
    }
    if (lane == 0) {
      if (a.row_min != nullptr) a.row_min[row] = mn;
      if (a.row_max != nullptr) a.row_max[row] = mx;
    }
  }
}

// dz = dr * mask / (1 - w) with dr the min-max backward of dout; per-CTA column partials of A_j = sum_b dr_bj * z_bj
// (the only batch reduction the gradient of w needs: d mask_j = A_j / (1-w_j), and through the division A_j mask_j / (1-w_j)^2).
template <int NV>
__global__ void __launch_bounds__(PG_THREADS, NV <= 6 ? 2 : 1) prigumbel_bwd_kernel(const PriGumbelBwdArgs a) {
  extern __shared__ float s_acc[];   // [PG_WARPS][H]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = static_cast<long long>(blockIdx.x) * PG_WARPS + warp; row < a.B;
       row += static_cast<long long>(gridDim.x) * PG_WARPS) {
    float4 z[NV], r[NV], g[NV];
    float mn, mx;
    int imn, imx;
    dropout_row<NV>(a.z + row * a.ldz, a.coef, a.H, lane, z, r, mn, mx, imn, imx);
    const float inv_r = 1.0f / (mx - mn);
    float s_min = 0.f, s_max = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int col = (lane + 32 * k) << 2;
      if (col < a.H) {
        g[k] = *reinterpret_cast<const float4*>(a.dout + row * a.ld_dout + col);
        const float e[4] = {r[k].x, r[k].y, r[k].z, r[k].w};
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
      if (col < a.H) {
        float dr[4] = {g[k].x * inv_r, g[k].y * inv_r, g[k].z * inv_r, g[k].w * inv_r};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (col + q == imn) dr[q] -= s_min;
          if (col + q == imx) dr[q] -= s_max;
        }
        const float4 m4 = *reinterpret_cast<const float4*>(a.coef + col);
        const float4 o4 = *reinterpret_cast<const float4*>(a.coef + a.H + col);
        const float4 i4 = *reinterpret_cast<const float4*>(a.coef + 3 * a.H + col);
        *reinterpret_cast<float4*>(a.dz + row * a.ld_dz + col) =
            make_float4(div_rn_z(dr[0] * m4.x, o4.x, i4.x), div_rn_z(dr[1] * m4.y, o4.y, i4.y),
                        div_rn_z(dr[2] * m4.z, o4.z, i4.z), div_rn_z(dr[3] * m4.w, o4.w, i4.w));
        acc[k].x = fmaf(dr[0], z[k].x, acc[k].x);
        acc[k].y = fmaf(dr[1], z[k].y, acc[k].y);
        acc[k].z = fmaf(dr[2], z[k].z, acc[k].z);
        acc[k].w = fmaf(dr[3], z[k].w, acc[k].w);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int col = (lane + 32 * k) << 2;
    if (col < a.H) *reinterpret_cast<float4*>(s_acc + warp * a.H + col) = acc[k];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < a.H; j += PG_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < PG_WARPS; ++wv) s += s_acc[wv * a.H + j];             // fixed order
    a.partial[static_cast<long long>(blockIdx.x) * a.H + j] = s;
  }
}

// dw_j = A_j * (dmask_j / (1-w_j) + mask_j / (1-w_j)^2) + [j == argmax] * (1 - e^eps) * wloss_scale
__global__ void __launch_bounds__(PG_THREADS) prigumbel_dw_kernel(const float* __restrict__ partial, int slabs, int H,
                                                                  const float* __restrict__ coef, const float* __restrict__ wloss,
                                                                  float exp_eps, float wloss_scale, float* __restrict__ dw,
                                                                  int accumulate) {
  const int j = blockIdx.x * PG_THREADS + threadIdx.x;
  if (j >= H) return;
  float A = 0.f;
  for (int s = 0; s < slabs; ++s) A += partial[static_cast<long long>(s) * H + j];
  const float mask = coef[j], omw = coef[H + j], dmask = coef[2 * H + j];
  float g = A * (dmask / omw + mask / (omw * omw));
  if (j == static_cast<int>(wloss[1])) g += (1.0f - exp_eps) * wloss_scale;
  dw[j] = accumulate ? dw[j] + g : g;
}

template <typename F>
int dispatch_nv(int H, F&& f) {
  const int nv = (H / 4 + 31) / 32;
  if (nv <= 6) return f(std::integral_constant<int, 6>{});
  if (nv <= 16) return f(std::integral_constant<int, 16>{});
  set_error("pgf_prigumbel: width H=%d exceeds the register-resident limit 2048 (the reference's is 768)", H);
  return PGF_ERR_UNSUPPORTED;
}

int row_grid(int B, int ctas_per_sm = 2) {
  long long g = (static_cast<long long>(B) + PG_WARPS - 1) / PG_WARPS;
  const long long cap = static_cast<long long>(ctas_per_sm) * num_sms();
  if (g > cap) g = cap;
  return g < 1 ? 1 : static_cast<int>(g);
}

}  // namespace

int prigumbel_coef(const float* w, const float* gum, int H, float exp_eps, float tau, int hard, unsigned long long seed,
                   unsigned int offset, float* coef, float* wloss, cudaStream_t s) {
  prigumbel_coef_kernel<<<1, PG_THREADS, 0, s>>>(w, gum, H, exp_eps, tau, hard, static_cast<unsigned int>(seed),
                                                 static_cast<unsigned int>(seed >> 32), offset, coef, wloss);
  PGF_CUDA_LAUNCH_CHECK("pgf_prigumbel_coef");
  return PGF_OK;
}

int prigumbel_fwd(const PriGumbelArgs& a, cudaStream_t s) {
  return dispatch_nv(a.H, [&](auto nv) {
    auto* k = &prigumbel_fwd_kernel<decltype(nv)::value>;
    const int occ = cached_occupancy(reinterpret_cast<const void*>(k), PG_THREADS, 0, 2);
    k<<<row_grid(a.B, occ), PG_THREADS, 0, s>>>(a);
    PGF_CUDA_LAUNCH_CHECK("pgf_prigumbel_fwd");
    return PGF_OK;
  });
}

int prigumbel_bwd_slabs(int B) { return row_grid(B); }

int prigumbel_bwd(const PriGumbelBwdArgs& a, float exp_eps, float wloss_scale, const float* wloss, float* dw, int accumulate,
                  cudaStream_t s) {
  const int grid = row_grid(a.B);
  const size_t smem = static_cast<size_t>(PG_WARPS) * a.H * sizeof(float);
  const int rc = dispatch_nv(a.H, [&](auto nv) {
    auto* k = &prigumbel_bwd_kernel<decltype(nv)::value>;
    ensure_dynamic_smem(reinterpret_cast<const void*>(k), smem);
    k<<<grid, PG_THREADS, smem, s>>>(a);
    PGF_CUDA_LAUNCH_CHECK("pgf_prigumbel_bwd");
    return PGF_OK;
  });
  if (rc != PGF_OK) return rc;
  if (dw != nullptr) {
    prigumbel_dw_kernel<<<(a.H + PG_THREADS - 1) / PG_THREADS, PG_THREADS, 0, s>>>(a.partial, grid, a.H, a.coef, wloss, exp_eps,
                                                                                   wloss_scale, dw, accumulate);
    PGF_CUDA_LAUNCH_CHECK("pgf_prigumbel_bwd(dw)");
  }
  return PGF_OK;
}

}  // namespace pgf
