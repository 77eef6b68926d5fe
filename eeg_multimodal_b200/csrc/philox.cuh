// Counter-based noise for the perturb/gate kernels: Philox4x32-10 keyed per
// (column group, global row, stream, offset).  Bit-for-bit definition: oracle/philox_ref.py.
// Replaces the reference's sequential global-RNG draws (models.py:74 Laplace.sample on the
// host + H2D copy; models.py:77 exponential_() inside F.gumbel_softmax).
#pragma once
#include <stdint.h>

namespace pgf {

// MUFU.LG2 without the denormal-input fix-up sequence (inputs here are >= 2^-24)
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#define PGF_STREAM_LAPLACE 0u
#define PGF_STREAM_GUMBEL0 1u
#define PGF_STREAM_GUMBEL1 2u

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1);

// key given as (k0,k1): the key schedule runs on the uniform datapath (used by the grouped launches,
// where the seed depends on blockIdx)
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += W0;
    k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Round keys of one Philox stream (key schedule k + r*W), computed on the HOST and passed inside the
// kernel-argument struct: indexed with compile-time constants they become constant-bank operands of
// the XORs, so the per-call key schedule costs no instructions at all.
struct PhiloxKeys {
  uint32_t k[20];
};
inline PhiloxKeys philox_make_keys(unsigned long long seed) {
  PhiloxKeys r;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    r.k[2 * i] = k0;
    r.k[2 * i + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return r;
}

// one round, pinned to 2 x IMAD.WIDE + 2 x LOP3
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
  uint32_t n0, n1, n2, n3;
  asm("{\n\t.reg .b64 p0, p1;\n\t.reg .b32 h0, l0, h1, l1;\n\t"
      "mul.wide.u32 p0, %4, 0xD2511F53;\n\t"
      "mul.wide.u32 p1, %6, 0xCD9E8D57;\n\t"
      "mov.b64 {l0, h0}, p0;\n\t"
      "mov.b64 {l1, h1}, p1;\n\t"
      "xor.b32 h1, h1, %5;\n\t"
      "xor.b32 %0, h1, %8;\n\t"
      "mov.b32 %1, l1;\n\t"
      "xor.b32 h0, h0, %7;\n\t"
      "xor.b32 %2, h0, %9;\n\t"
      "mov.b32 %3, l0;\n\t}"
      : "=r"(n0), "=r"(n1), "=r"(n2), "=r"(n3)
      : "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(k0), "r"(k1));
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__device__ __forceinline__ uint4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& rk) {
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, rk.k[2 * r], rk.k[2 * r + 1]);
  return make_uint4(c0, c1, c2, c3);
}

// Laplace(0,1): bit 31 = sign, bits 0..22 = magnitude uniform, v=(m+0.5)*2^-23 exact in fp32,
// x = -+ln(v).  `laplace_scaled_from_bits(r, c)` returns x * eps_hat given c = -ln2 * eps_hat:
// lg2(v) * c with the sign bit XOR-ed in (5 instructions + 1 MUFU).
__device__ __forceinline__ float laplace_scaled_from_bits(uint32_t r, float c) {
  // (m + 0.5) * 2^-23 without an int->float conversion (I2F shares the quarter-rate pipe with MUFU):
  // 1.m as a float in [1,2), minus (1 - 2^-24); the result (2m+1)*2^-24 is exact.
  uint32_t one_m;  // (r & 0x7FFFFF) | 0x3F800000 as ONE three-input logic op (both constants in registers)
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(one_m) : "r"(r), "r"(0x7FFFFFu), "r"(0x3F800000u));
  const float v = __uint_as_float(one_m) - 0.99999994039535522f;
  uint32_t cs;     // c with the sign bit of r XOR-ed in: c ^ (r & 0x80000000)
  asm("lop3.b32 %0, %1, %2, %3, 0x78;" : "=r"(cs) : "r"(__float_as_uint(c)), "r"(r), "r"(0x80000000u));
  return lg2_ftz(v) * __uint_as_float(cs);
}
// Perturbed feature as ONE fused multiply-add on top of the normalised value xn:
//   xn + eps_hat * Laplace(r)  =  lg2(v) * (c ^ sign(r)) + xn,   c = -ln2 * eps_hat.
// Every Philox-mode forward kernel uses exactly this sequence, so grouped, single-model and TMA-ring launches
// are bit-identical.
__device__ __forceinline__ float perturb_fma(float xn, uint32_t r, float c) {
  uint32_t one_m, cs;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(one_m) : "r"(r), "r"(0x7FFFFFu), "r"(0x3F800000u));
  asm("lop3.b32 %0, %1, %2, %3, 0x78;" : "=r"(cs) : "r"(__float_as_uint(c)), "r"(r), "r"(0x80000000u));
  const float v = __uint_as_float(one_m) - 0.99999994039535522f;
  return __fmaf_rn(lg2_ftz(v), __uint_as_float(cs), xn);
}
// lg2(v) with the Laplace sign of r folded into its sign bit: perturb_fma(xn, r, c) == fma(signed_lg2_from_bits(r), c, xn)
// bit for bit (flipping the sign of either factor of a product gives the same result)
__device__ __forceinline__ float signed_lg2_from_bits(uint32_t r) {
  uint32_t one_m;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(one_m) : "r"(r), "r"(0x7FFFFFu), "r"(0x3F800000u));
  const float l = lg2_ftz(__uint_as_float(one_m) - 0.99999994039535522f);
  return __uint_as_float(__float_as_uint(l) ^ (r & 0x80000000u));
}
__device__ __forceinline__ float laplace_from_bits(uint32_t r) {
  return laplace_scaled_from_bits(r, -0.69314718055994531f);
}

// Gumbel(0,1) = -log(Exp(1)), Exp(1) = -log(v), v=((r>>9)+0.5)*2^-23.
__device__ __forceinline__ float gumbel_from_bits(uint32_t r) {
  const float v = __uint_as_float(0x3F800000u | (r >> 9)) - 0.99999994039535522f;  // ((r>>9)+0.5)*2^-23, exact
  return -__logf(-logf(v));  // accurate inner log: E = -ln(v) can be tiny
}

}  // namespace pgf
