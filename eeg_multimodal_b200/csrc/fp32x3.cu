// fp32-parity route of the large-batch fusion MLP onto the tensor cores.
//
// Reference arithmetic: nn.Linear in fp32 (python/src/custom_models/models.py:46-51,80).  The tcgen05 bf16 GEMM keeps
// 8 significand bits per operand; north_star's fp32 bar (1e-5 on logits and gradients) needs all 24.  Every fp32
// value is therefore written as three bf16 planes
//     hi = bf16(x)   mid = bf16(x - hi)   lo = bf16(x - hi - mid)          hi + mid + lo == x   (exactly)
// (round-to-nearest each time: the remainders are exact in fp32 and fit the planes that are left, so nothing is lost
// above the bf16 exponent floor), and pgf_gemm_bf16x3 (gemm_tc.cu, nseg = 6) contracts the six plane pairs whose
// products reach down to 2^-24 of the result inside one fp32 accumulation.
//
// split3_kernel is the elementwise stage between two such GEMMs: it applies what the reference applies between two
// nn.Linear calls -- bias, ReLU / Tanh (libm tanhf, not the 3e-7 approximation of the bf16 epilogues), or the ReLU
// derivative mask of the backward pass -- and emits the planes the next GEMM reads and/or the fp32 tensor itself.
#include "pgf_kernels.cuh"

#define PGF_ACT_NONE 0
#define PGF_ACT_RELU 1
#define PGF_ACT_TANH 2

namespace pgf {

__device__ __forceinline__ void split3_one(float v, __nv_bfloat16& hi, __nv_bfloat16& mid, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(hi);   // exact
  mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);  // exact
  lo = __float2bfloat16_rn(r2);
}

// One thread per 8 consecutive columns of a row: two 128-bit loads, one 128-bit store per output plane.
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ src, long long ld, int R, int C8,
                                                     const float* __restrict__ bias, int act,
                                                     const __nv_bfloat16* __restrict__ mask, long long ld_mask,
                                                     float* __restrict__ out_f32, long long ld_out,
                                                     __nv_bfloat16* __restrict__ planes, long long ldp, long long plane_stride) {
  const long long total = static_cast<long long>(R) * C8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / C8), c = static_cast<int>(i - static_cast<long long>(r) * C8) * 8;
    const float4* sp = reinterpret_cast<const float4*>(src + static_cast<long long>(r) * ld + c);
    const float4 a = sp[0], b = sp[1];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (bias) {
      const float4 ba = __ldg(reinterpret_cast<const float4*>(bias + c)), bb = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
      v[0] += ba.x; v[1] += ba.y; v[2] += ba.z; v[3] += ba.w;
      v[4] += bb.x; v[5] += bb.y; v[6] += bb.z; v[7] += bb.w;
    }
    if (act == PGF_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : (v[j] != v[j] ? v[j] : 0.f);  // torch.relu keeps NaN
    } else if (act == PGF_ACT_TANH) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tanhf(v[j]);
    }
    if (mask) {  // ReLU backward: the hi plane of the activation has the activation's sign (hi == 0 iff x == 0)
      const uint4 m = *reinterpret_cast<const uint4*>(mask + static_cast<long long>(r) * ld_mask + c);
      const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(mw[j]);
        v[2 * j] = f.x > 0.f ? v[2 * j] : 0.f;
        v[2 * j + 1] = f.y > 0.f ? v[2 * j + 1] : 0.f;
      }
    }
    if (out_f32) {
      float4* op = reinterpret_cast<float4*>(out_f32 + static_cast<long long>(r) * ld_out + c);
      op[0] = make_float4(v[0], v[1], v[2], v[3]);
      op[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (planes) {
      __nv_bfloat16 h[8], m[8], l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split3_one(v[j], h[j], m[j], l[j]);
      __nv_bfloat16* pp = planes + static_cast<long long>(r) * ldp + c;
      *reinterpret_cast<uint4*>(pp) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(pp + plane_stride) = *reinterpret_cast<const uint4*>(m);
      *reinterpret_cast<uint4*>(pp + 2 * plane_stride) = *reinterpret_cast<const uint4*>(l);
    }
  }
}

int split3(const float* src, long long ld, int R, int C, const float* bias, int act, const void* mask_plane, long long ld_mask,
           float* out_f32, long long ld_out, void* planes, long long ldp, long long plane_stride, cudaStream_t s) {
  const long long items = static_cast<long long>(R) * (C / 8);
  long long blocks = (items + 255) / 256;
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  split3_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(src, ld, R, C / 8, bias, act,
                                                               static_cast<const __nv_bfloat16*>(mask_plane), ld_mask, out_f32,
                                                               ld_out, static_cast<__nv_bfloat16*>(planes), ldp, plane_stride);
  PGF_CUDA_LAUNCH_CHECK("pgf_split3");
  return PGF_OK;
}

}  // namespace pgf
