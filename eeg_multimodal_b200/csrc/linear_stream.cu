// fp32 CUDA-core path of the fusion MLP at the reference's batch size (kernel (b), small-batch / parity regime):
// persistent, TMA-fed weight-streaming kernels.
//
// Reference ops replaced: nn.Linear + ReLU / Tanh of `fc_layers`, their autograd, and model_optimizer.step() on them
// (models.py:46-51,80; past_acc.py:87-92,137,211-212).  At B=8 a layer is [8,K]x[K,N]: ~8 flop per weight byte, bound by
// streaming the fp32 weights from HBM.  Round 1 streamed the weights through registers (two 128-bit loads in flight per
// lane): with the few models one GPU of eight owns (6), every launch is ONE wave whose CTAs all stage their activation tile
// and then all stream, and the bytes in flight per SM are capped by the register file -- ncu: 0.53 / 0.41 / 0.18 / 0.39 of
// the HBM rate for fwd(W1) / fwd(W2) / dX(W2) / dX(W1).  Here every kernel is
//   * persistent: one CTA per SM, a contiguous range of work items each (grid = SM count, no tail wave);
//   * fed by the TMA engine: one elected producer thread streams [32 rows x 128 floats] weight boxes (3-D tensor map over
//     [model][row][k]) into a shared-memory ring, 6-8 stages = 96-128 KB in flight per SM at no register cost; the eight
//     consumer warps (4 rows of a box each) wait on mbarriers, so nothing stalls on a scoreboard;
//   * grouped over the models of a sweep through the tensor map's third dimension.
//   fwd : Y[b,n]  = act(sum_k X[b,k] W[n,k] + bias[n])      item = (model, batch chunk, 32-row block), all of K
//   dx  : dX[b,k] = sum_n dY[b,n] W[n,k]  (* act'(mask))     item = (model, batch chunk, row split, 128-column chunk)
//   adam: W,m,v <- Adam(dW = dY^T X)  (+ bias)               item = (layer, model, 128-column chunk, row split); the rank-8
//         gradient is recomputed per element instead of being written and read back (24 instead of 32 B per parameter)
// Accumulation is fp32 FMA in a fixed order (deterministic; no atomics on data).  fwd and adam keep round 1's per-element
// operation order bit for bit; dx sums its rows per warp, then warps 0..7, then row splits 0..n in that order.
#include <cuda.h>
#include <stdlib.h>

#include "pgf_kernels.cuh"

namespace pgf {

#define PGF_ACT_NONE 0
#define PGF_ACT_RELU 1
#define PGF_ACT_TANH 2

constexpr int TB = 8;                         // batch rows per tile (the reference's batch size)
constexpr int LS_ROWS = 32;                   // weight rows per ring stage
constexpr int LS_KC = 128;                    // floats per row chunk (512 B)
constexpr int LS_BOX = LS_ROWS * LS_KC;       // floats per box (16 KB)
constexpr int LS_WARPS = 8;                   // consumer warps: 4 rows of a box each
constexpr int LS_CONSUMERS = LS_WARPS * 32;
constexpr int LS_THREADS = LS_CONSUMERS + 32; // + one producer warp (one elected lane issues every copy)
constexpr size_t LS_SMEM_MAX = 220 * 1024;

// ---- mbarrier / TMA helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ls_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ls_bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ls_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ls_bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = ls_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void ls_bar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ls_u32(bar)) : "memory");
}
__device__ __forceinline__ void ls_bar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ls_u32(bar)), "r"(bytes) : "memory");
}
// [32 rows x 128 floats] box at (k0, n0, model) of a [model][row][k] tensor -> shared memory, completion on `bar`
__device__ __forceinline__ void ls_tma_box(float* dst, const CUtensorMap* tm, int k0, int n0, int model, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          ls_u32(dst)),
      "l"(tm), "r"(k0), "r"(n0), "r"(model), "r"(ls_u32(bar))
      : "memory");
}
// The same load with an L2 eviction-priority hint.  A weight matrix small enough to live in the 126 MB L2 next to the
// streams that pass through it (fc_layers.2 of the 6 models one GPU of eight owns: 42 MB) is read five times per step --
// two forward passes, two dX passes, the Adam update -- and only the first read has to come from HBM if its lines are
// loaded `evict_last` while everything that streams (fc_layers.0, the Adam moments) keeps the normal priority.
__device__ __forceinline__ uint64_t ls_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void ls_tma_box_hint(float* dst, const CUtensorMap* tm, int k0, int n0, int model, uint64_t* bar,
                                                uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], "
      "[%5], %6;" ::"r"(ls_u32(dst)),
      "l"(tm), "r"(k0), "r"(n0), "r"(model), "r"(ls_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void ls_tma_w(float* dst, const CUtensorMap* tm, int k0, int n0, int model, uint64_t* bar, bool keep,
                                         uint64_t policy) {
  if (keep) ls_tma_box_hint(dst, tm, k0, n0, model, bar, policy);
  else ls_tma_box(dst, tm, k0, n0, model, bar);
}
// contiguous bytes -> shared memory (activation / gradient rows)
__device__ __forceinline__ void ls_bulk_row(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ls_u32(dst)),
               "l"(src), "r"(bytes), "r"(ls_u32(bar))
               : "memory");
}
// Rows of an activation / gradient tile into shared memory, signalled on `bar` (count 1).  16-byte aligned rows go through
// the TMA engine; anything else (odd widths in tests) is copied by the calling thread itself and published by a plain arrive.
__device__ __forceinline__ void ls_load_rows(float* dst, long long dst_ld, const float* src, long long src_ld, int rows, int n,
                                             uint64_t* bar) {
  const bool bulk = ((n & 3) == 0) && ((src_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  if (bulk) {
    ls_bar_expect(bar, static_cast<uint32_t>(rows) * static_cast<uint32_t>(n) * 4u);
    for (int b = 0; b < rows; ++b) ls_bulk_row(dst + b * dst_ld, src + b * src_ld, static_cast<uint32_t>(n) * 4u, bar);
  } else {
    for (int b = 0; b < rows; ++b)
      for (int i = 0; i < n; ++i) dst[b * dst_ld + i] = src[b * src_ld + i];
    ls_bar_arrive(bar);
  }
}
__device__ __forceinline__ void ls_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(LS_CONSUMERS) : "memory"); }

struct LsRingPos {
  int stage;
  uint32_t phase;
  __device__ __forceinline__ void next(int n_stages) {
    if (++stage == n_stages) {
      stage = 0;
      phase ^= 1u;
    }
  }
};

// contiguous item range of this CTA
__device__ __forceinline__ void ls_item_range(int n_items, int& lo, int& hi) {
  lo = static_cast<int>(static_cast<long long>(n_items) * blockIdx.x / gridDim.x);
  hi = static_cast<int>(static_cast<long long>(n_items) * (blockIdx.x + 1) / gridDim.x);
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
struct LsFwdArgs {
  const float* bias; long long sb;
  float* Y; long long ldy, sY;
  int B, N, K, act, n_models;
  int row_blocks, bchunks, kchunks, stages;
  int x_models;                  // 1: the activations are shared by all models (model coordinate 0 of the X map)
  int dbg;                       // probe (PGF_LS_DBG=1): consumers skip the arithmetic (pure TMA delivery rate)
  int w_keep;                    // the weights of all models fit L2: load them evict_last (ls_policy_evict_last)
};
constexpr int LS_XBOX = TB * LS_KC;            // floats of an activation chunk [8][128]
constexpr int LS_FSTAGE = LS_BOX + LS_XBOX;    // forward ring stage: weight box + the activation chunk it multiplies (20 KB)

// A stage carries the [32 x 128] weight box AND the [8 x 128] activation chunk it meets (second tensor map over
// [model][batch row][k], rows past B zero-filled): no separate activation tile has to land before the first box can be
// consumed (that tile was 72 KB = 2.5 us of every launch), at the price of re-fetching 4 KB of L2-resident activations per
// 16 KB of weights.
__global__ void __launch_bounds__(LS_THREADS, 1) ls_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                                                                const LsFwdArgs a) {
  extern __shared__ __align__(1024) unsigned char ls_smem[];
  float* ring = reinterpret_cast<float*>(ls_smem);                         // [stages][ W 32x128 | X 8x128 ]
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(a.stages) * LS_FSTAGE);
  uint64_t* empty = full + a.stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ls_bar_init(full + s, 1);
      ls_bar_init(empty + s, LS_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  griddep_wait();
  griddep_launch();
  const int n_items = a.n_models * a.bchunks * a.row_blocks;
  int lo, hi;
  ls_item_range(n_items, lo, hi);
  if (warp == LS_WARPS) {
    // ---- producer: per item, its weight boxes along K, each with the matching activation chunk
    if (lane != 0) return;
    LsRingPos pos{0, 0};
    const uint64_t keep_policy = ls_policy_evict_last();
    for (int item = lo; item < hi; ++item) {
      const int rb = item % a.row_blocks, key = item / a.row_blocks;
      const int bc = key % a.bchunks, model = key / a.bchunks;
      for (int c = 0; c < a.kchunks; ++c) {
        ls_bar_wait(empty + pos.stage, pos.phase ^ 1u);
        ls_bar_expect(full + pos.stage, LS_FSTAGE * 4u);
        float* dst = ring + static_cast<size_t>(pos.stage) * LS_FSTAGE;
        ls_tma_w(dst, &tmW, c * LS_KC, rb * LS_ROWS, model, full + pos.stage, a.w_keep != 0, keep_policy);
        ls_tma_box(dst + LS_BOX, &tmX, c * LS_KC, bc * TB, a.x_models > 1 ? model : 0, full + pos.stage);
        pos.next(a.stages);
      }
    }
    return;
  }
  // ---- consumers: warp w owns rows 4w..4w+3 of every box; lane l owns the l-th float4 of a 128-float chunk
  constexpr int R = LS_ROWS / LS_WARPS;
  LsRingPos pos{0, 0};
  for (int item = lo; item < hi; ++item) {
    const int rb = item % a.row_blocks, key = item / a.row_blocks;
    const int bc = key % a.bchunks, model = key / a.bchunks;
    float acc[R * TB];
#pragma unroll
    for (int i = 0; i < R * TB; ++i) acc[i] = 0.f;
    for (int c = 0; c < a.kchunks; ++c) {
      ls_bar_wait(full + pos.stage, pos.phase);
      const float* st = ring + static_cast<size_t>(pos.stage) * LS_FSTAGE;
      const float4* box = reinterpret_cast<const float4*>(st) + (warp * R) * (LS_KC / 4) + lane;
      const float4* xs = reinterpret_cast<const float4*>(st + LS_BOX) + lane;
      float4 w[R];
#pragma unroll
      for (int r = 0; r < R; ++r) w[r] = box[r * (LS_KC / 4)];
      if (!(a.dbg & 1)) {   // past K (and past B) both boxes are zero-filled: those products add exact zeros
#pragma unroll
        for (int b = 0; b < TB; ++b) {
          const float4 x = xs[b * (LS_KC / 4)];
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
      }
      __syncwarp();
      if (lane == 0) ls_bar_arrive(empty + pos.stage);
      pos.next(a.stages);
    }
    // warp reduction of R*TB values; lane i < R*TB writes value i (i = r*TB + b)
#pragma unroll
    for (int i = 0; i < R * TB; ++i) acc[i] = warp_sum(acc[i]);
    float mine = 0.f;
#pragma unroll
    for (int i = 0; i < R * TB; ++i)
      if (lane == i) mine = acc[i];
    const int nb = min(TB, a.B - bc * TB);
    const int r = lane / TB, b = lane - r * TB;
    const int n = rb * LS_ROWS + warp * R + r;
    if (n < a.N && b < nb) {
      float v = mine + (a.bias ? a.bias[model * a.sb + n] : 0.f);
      if (a.act == PGF_ACT_RELU) v = fmaxf(v, 0.f);
      else if (a.act == PGF_ACT_TANH) v = tanhf(v);
      a.Y[model * a.sY + static_cast<long long>(bc * TB + b) * a.ldy + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dX
// ------------------------------------------------------------------------------------------------------------------
struct LsDxArgs {
  const float* dY; long long ldy, sdY;
  const float* mask_src; int mask_mode; long long ld_mask, s_mask;
  float* dX; long long ldx, sdX;
  float* partial;            // [n_models][bchunks][kchunks][nsplit][TB][128]   (nsplit > 1)
  unsigned int* counters;    // [n_models][bchunks][kchunks], zero on entry, zero again on exit
  int B, N, K, n_models;
  int kchunks, bchunks, nsplit, rows_per_split, stages;
  int w_keep;                // see LsFwdArgs
};

__device__ __forceinline__ float4 ls_dx_mask(float4 s, const float* mask_src, int mask_mode, long long moff) {
  if (mask_src) {
    const float4 m = *reinterpret_cast<const float4*>(mask_src + moff);
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
  return s;
}

__global__ void __launch_bounds__(LS_THREADS, 1) ls_dx_kernel(const __grid_constant__ CUtensorMap tmW, const LsDxArgs a) {
  extern __shared__ __align__(1024) unsigned char ls_smem[];
  __shared__ int s_last;
  float* ring = reinterpret_cast<float*>(ls_smem);                         // [stages][32][128]
  float* red = ring + static_cast<size_t>(a.stages) * LS_BOX;              // [8 warps][8][128] cross-warp reduction
  float* sdy2 = red + LS_WARPS * TB * LS_KC;                               // [2][8][rows_per_split] gradient tiles (batch-major),
  const size_t tile_floats = static_cast<size_t>(TB) * a.rows_per_split;   // double-buffered: a CTA's items are `grid` apart
  uint64_t* full = reinterpret_cast<uint64_t*>(sdy2 + 2 * tile_floats);
  uint64_t* empty = full + a.stages;
  uint64_t* xfull = empty + a.stages;      // [2]
  uint64_t* xempty = xfull + 2;            // [2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ls_bar_init(full + s, 1);
      ls_bar_init(empty + s, LS_WARPS);
    }
    for (int j = 0; j < 2; ++j) {
      ls_bar_init(xfull + j, 1);
      ls_bar_init(xempty + j, LS_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  griddep_wait();
  griddep_launch();
  // contiguous item ranges, the column chunk running fastest: consecutive items of a CTA share their gradient tile
  // (dealing the items round-robin, as the gradient+Adam kernel does for the sake of its stores, cost this read-only
  // kernel 2-3 us per launch in tile reloads)
  const int n_items = a.n_models * a.bchunks * a.nsplit * a.kchunks;
  const int RS = a.rows_per_split;
  int lo, hi;
  ls_item_range(n_items, lo, hi);
  if (warp == LS_WARPS) {
    if (lane != 0) return;
    LsRingPos pos{0, 0};
    const uint64_t keep_policy = ls_policy_evict_last();
    int ts = -1, xkey = -1;                  // tile sequence number: buffer ts & 1, used (ts >> 1) times before
    for (int item = lo; item < hi; ++item) {
      const int c = item % a.kchunks, key = item / a.kchunks;
      const int sp = key % a.nsplit, t = key / a.nsplit;
      const int bc = t % a.bchunks, model = t / a.bchunks;
      const int n0 = sp * RS, n1 = min(a.N, n0 + RS);
      if (key != xkey) {
        xkey = key;
        ++ts;
        const int j = ts & 1, u = ts >> 1;
        if (u > 0) ls_bar_wait(xempty + j, static_cast<uint32_t>(u - 1) & 1u);
        const int nb = min(TB, a.B - bc * TB);
        ls_load_rows(sdy2 + j * tile_floats, RS, a.dY + model * a.sdY + static_cast<long long>(bc) * TB * a.ldy + n0, a.ldy, nb, n1 - n0,
                     xfull + j);
      }
      for (int n = n0; n < n1; n += LS_ROWS) {
        ls_bar_wait(empty + pos.stage, pos.phase ^ 1u);
        ls_bar_expect(full + pos.stage, LS_BOX * 4u);
        ls_tma_w(ring + static_cast<size_t>(pos.stage) * LS_BOX, &tmW, c * LS_KC, n, model, full + pos.stage, a.w_keep != 0, keep_policy);
        pos.next(a.stages);
      }
    }
    return;
  }
  constexpr int R = LS_ROWS / LS_WARPS;
  LsRingPos pos{0, 0};
  int ts = -1, xkey = -1;
  const float* sdy = sdy2;
  for (int item = lo; item < hi; ++item) {
    const int c = item % a.kchunks, key = item / a.kchunks;
    const int sp = key % a.nsplit, t = key / a.nsplit;
    const int bc = t % a.bchunks, model = t / a.bchunks;
    const int n0 = sp * RS, n1 = min(a.N, n0 + RS);
    const int nb = min(TB, a.B - bc * TB);
    if (key != xkey) {
      xkey = key;
      ++ts;
      sdy = sdy2 + (ts & 1) * tile_floats;
      ls_bar_wait(xfull + (ts & 1), static_cast<uint32_t>(ts >> 1) & 1u);
    }
    float4 acc[TB];
#pragma unroll
    for (int b = 0; b < TB; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = n0; n < n1; n += LS_ROWS) {
      ls_bar_wait(full + pos.stage, pos.phase);
      const float4* box = reinterpret_cast<const float4*>(ring + static_cast<size_t>(pos.stage) * LS_BOX) + (warp * R) * (LS_KC / 4) + lane;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int nl = n - n0 + warp * R + r;       // row inside the split; rows past N are zero boxes with no gradient behind them
        if (n + warp * R + r < n1) {
          const float4 w = box[r * (LS_KC / 4)];
#pragma unroll
          for (int b = 0; b < TB; ++b) {
            const float g = b < nb ? sdy[static_cast<size_t>(b) * RS + nl] : 0.f;
            acc[b].x = fmaf(g, w.x, acc[b].x);
            acc[b].y = fmaf(g, w.y, acc[b].y);
            acc[b].z = fmaf(g, w.z, acc[b].z);
            acc[b].w = fmaf(g, w.w, acc[b].w);
          }
        }
      }
      __syncwarp();
      if (lane == 0) ls_bar_arrive(empty + pos.stage);
      pos.next(a.stages);
    }
    if (item + 1 >= hi || (item + 1) / a.kchunks != key) {   // last item on this tile
      __syncwarp();
      if (lane == 0) ls_bar_arrive(xempty + (ts & 1));
    }
    // ---- warps 0..7 in order, then (nsplit > 1) the row splits in order
#pragma unroll
    for (int b = 0; b < TB; ++b) *reinterpret_cast<float4*>(red + (static_cast<size_t>(warp) * TB + b) * LS_KC + lane * 4) = acc[b];
    ls_consumer_sync();
    const int ob = warp;                              // this thread's output: batch row `warp`, float4 `lane` of the chunk
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < LS_WARPS; ++w) {
      const float4 p = *reinterpret_cast<const float4*>(red + (static_cast<size_t>(w) * TB + ob) * LS_KC + lane * 4);
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    const int kk = c * LS_KC + lane * 4;
    const long long row = static_cast<long long>(bc) * TB + ob;
    const bool valid = kk < a.K && ob < nb;
    if (a.nsplit == 1) {
      if (valid)
        *reinterpret_cast<float4*>(a.dX + model * a.sdX + row * a.ldx + kk) =
            ls_dx_mask(s, a.mask_src, a.mask_mode, model * a.s_mask + row * a.ld_mask + kk);
      ls_consumer_sync();                             // `red` is rewritten by the next item
      continue;
    }
    const long long unit = (static_cast<long long>(model) * a.bchunks + bc) * a.kchunks + c;
    float4* P = reinterpret_cast<float4*>(a.partial) + (unit * a.nsplit * TB) * (LS_KC / 4);
    P[(static_cast<long long>(sp) * TB + ob) * (LS_KC / 4) + lane] = s;
    __threadfence();
    ls_consumer_sync();
    if (threadIdx.x == 0) s_last = atomicAdd(a.counters + unit, 1u) == static_cast<unsigned int>(a.nsplit) - 1u;
    ls_consumer_sync();
    if (s_last) {
      __threadfence();
      if (threadIdx.x == 0) a.counters[unit] = 0u;
      float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = 0; i < a.nsplit; ++i) {
        const float4 p = __ldcg(P + (static_cast<long long>(i) * TB + ob) * (LS_KC / 4) + lane);
        tot.x += p.x; tot.y += p.y; tot.z += p.z; tot.w += p.w;
      }
      if (valid)
        *reinterpret_cast<float4*>(a.dX + model * a.sdX + row * a.ldx + kk) =
            ls_dx_mask(tot, a.mask_src, a.mask_mode, model * a.s_mask + row * a.ld_mask + kk);
    }
    ls_consumer_sync();                               // s_last / `red` are rewritten by the next item
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dW fused into Adam: dW[n,k] = sum_b dY[b,n] X[b,k] is a rank-B outer product, so at the reference batch size it is
// recomputed inside the optimiser (8 FMAs per element) instead of being written to HBM by one kernel and read back by
// the next.  W, m and v arrive through three tensor maps of the same geometry; the update is written straight from
// registers.  The arithmetic per element is adam_update() on the gradient summed in batch order.
// ------------------------------------------------------------------------------------------------------------------
struct LsAdamLayerDev {
  const float* dY; long long ldy, sdY;
  const float* X; long long ldx, sX;
  float* W; float* mW; float* vW;
  float* bias; float* mb; float* vb;
  int N, K, kchunks, rsplit, rows_per_split, items;
  int w_keep;                // see LsFwdArgs (the weight loads only: the moments stream)
};
struct LsAdamArgs {
  LsAdamLayerDev l[2];
  int n_layers, n_models, B, stages, rows_max;
  long long sP;
  AdamCoef c;
  const StepState* st;
  StepAdvance adv;
  int dbg;   // probes (PGF_LS_DBG): 1 = consumers do not store, 2 = consumers skip the update arithmetic
};
struct LsAdamMaps {
  CUtensorMap w[2], m[2], v[2];
};

constexpr int LA_WARPS = 16;                  // consumer warps of the gradient+Adam kernel: 2 rows of a box each (the IEEE sqrt /
constexpr int LA_CONSUMERS = LA_WARPS * 32;   // division chains of the update need more warps in flight than the FMA kernels)
constexpr int LA_THREADS = LA_CONSUMERS + 32;

__global__ void __launch_bounds__(LA_THREADS, 1) ls_adam_kernel(const __grid_constant__ LsAdamMaps tm, const LsAdamArgs a) {
  extern __shared__ __align__(1024) unsigned char ls_smem[];
  float* ring = reinterpret_cast<float*>(ls_smem);                         // [stages][3][32][128]  (W, m, v boxes)
  float* sdy2 = ring + static_cast<size_t>(a.stages) * 3 * LS_BOX;         // [2][8][rows_max] gradient tiles (batch-major),
  const size_t tile_floats = static_cast<size_t>(TB) * a.rows_max;         // double-buffered: the tile changes with every item
  float* xs2 = sdy2 + 2 * tile_floats;                                     // [2][8][128] the item's columns of the layer input
  uint64_t* full = reinterpret_cast<uint64_t*>(xs2 + 2 * TB * LS_KC);
  uint64_t* empty = full + a.stages;
  uint64_t* xfull = empty + a.stages;      // [2]
  uint64_t* xempty = xfull + 2;            // [2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      ls_bar_init(full + s, 1);
      ls_bar_init(empty + s, LA_WARPS);
    }
    for (int j = 0; j < 2; ++j) {
      ls_bar_init(xfull + j, 2);    // two tile loads (gradient rows, input columns) arrive once each
      ls_bar_init(xempty + j, LA_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  griddep_wait();
  griddep_launch();
  const int n_items = a.l[0].items + (a.n_layers > 1 ? a.l[1].items : 0);
  // item -> (layer, model, row split, column chunk), the column chunk running fastest, items dealt round-robin over the
  // CTAs: neighbouring CTAs read AND write neighbouring 512-byte chunks of the same rows of W, m and v at the same time,
  // so open DRAM pages are used completely (with every CTA in its own row range the stores alone cost 107 us of a 211 us
  // launch: 5.0 TB/s for a read+write stream that the copy engine moves at 6.5)
#define LS_ADAM_DECODE(item)                                                               \
  const int layer = ((item) >= a.l[0].items) ? 1 : 0;                                      \
  const LsAdamLayerDev& L = a.l[layer];                                                    \
  const int li = (item) - (layer ? a.l[0].items : 0);                                      \
  const int c = li % L.kchunks, t_ = li / L.kchunks;                                       \
  const int sp = t_ % L.rsplit, model = t_ / L.rsplit;                                     \
  const int n0 = sp * L.rows_per_split, n1 = min(L.N, n0 + L.rows_per_split);
  if (warp == LA_WARPS) {
    if (lane != 0) return;
    LsRingPos pos{0, 0};
    const uint64_t keep_policy = ls_policy_evict_last();
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      LS_ADAM_DECODE(item)
      const int j = it & 1, u = it >> 1;   // tile buffers and how often they were used before
      if (u > 0) ls_bar_wait(xempty + j, static_cast<uint32_t>(u - 1) & 1u);
      float* sdy = sdy2 + j * tile_floats;
      ls_load_rows(sdy, a.rows_max, L.dY + model * L.sdY + n0, L.ldy, a.B, n1 - n0, xfull + j);
      ls_load_rows(xs2 + j * TB * LS_KC, LS_KC, L.X + model * L.sX + c * LS_KC, L.ldx, a.B, min(LS_KC, L.K - c * LS_KC), xfull + j);
      const CUtensorMap* mw = layer ? &tm.w[1] : &tm.w[0];
      const CUtensorMap* mm = layer ? &tm.m[1] : &tm.m[0];
      const CUtensorMap* mv = layer ? &tm.v[1] : &tm.v[0];
      for (int n = n0; n < n1; n += LS_ROWS) {
        ls_bar_wait(empty + pos.stage, pos.phase ^ 1u);
        ls_bar_expect(full + pos.stage, 3u * LS_BOX * 4u);
        float* dst = ring + static_cast<size_t>(pos.stage) * 3 * LS_BOX;
        ls_tma_w(dst, mw, c * LS_KC, n, model, full + pos.stage, L.w_keep != 0, keep_policy);
        ls_tma_box(dst + LS_BOX, mm, c * LS_KC, n, model, full + pos.stage);
        ls_tma_box(dst + 2 * LS_BOX, mv, c * LS_KC, n, model, full + pos.stage);
        pos.next(a.stages);
      }
    }
    return;
  }
  constexpr int R = LS_ROWS / LA_WARPS;
  const AdamCoef coef = adam_coef_at(a.c, a.st, 1);
  LsRingPos pos{0, 0};
  int it = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
    LS_ADAM_DECODE(item)
    const int j = it & 1, u = it >> 1;
    const float* sdy = sdy2 + j * tile_floats;
    const float* xs = xs2 + j * TB * LS_KC;
    const int kk = c * LS_KC + lane * 4;
    const bool kvalid = kk < L.K;
    ls_bar_wait(xfull + j, static_cast<uint32_t>(u) & 1u);
    if (c == 0 && L.bias) {   // bias: gradient = column sum of dY, one thread per row of the item
      for (int nl = threadIdx.x; nl < n1 - n0; nl += LA_CONSUMERS) {
        float g = 0.f;
#pragma unroll
        for (int b = 0; b < TB; ++b) g += b < a.B ? sdy[static_cast<size_t>(b) * a.rows_max + nl] : 0.f;
        const long long i = model * a.sP + n0 + nl;
        float p = L.bias[i], m = L.mb[i], v = L.vb[i];
        adam_update(p, m, v, g, coef);
        L.bias[i] = p; L.mb[i] = m; L.vb[i] = v;
      }
    }
    for (int n = n0; n < n1; n += LS_ROWS) {
      ls_bar_wait(full + pos.stage, pos.phase);
      const float* st = ring + static_cast<size_t>(pos.stage) * 3 * LS_BOX;
      // both rows of this warp as straight-line code (loads, 8 + 8 gradient FMAs, 8 independent updates), stores predicated:
      // the IEEE sqrt / division chains of the updates overlap instead of running one after the other
      float4 p[R], m[R], v[R], g[R];
      bool ok[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = n + warp * R + r;
        ok[r] = row < n1 && kvalid;
        const int o = (warp * R + r) * LS_KC + lane * 4;
        p[r] = *reinterpret_cast<const float4*>(st + o);
        m[r] = *reinterpret_cast<const float4*>(st + LS_BOX + o);
        v[r] = *reinterpret_cast<const float4*>(st + 2 * LS_BOX + o);
        g[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int b = 0; b < TB; ++b) {   // batch order: the gradient of linear_dw_kernel, bit for bit
        const float4 xb = (b < a.B && kvalid) ? *reinterpret_cast<const float4*>(xs + b * LS_KC + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int nl = min(n + warp * R + r, n1 - 1) - n0;
          const float gy = b < a.B ? sdy[static_cast<size_t>(b) * a.rows_max + nl] : 0.f;
          g[r].x = fmaf(gy, xb.x, g[r].x);
          g[r].y = fmaf(gy, xb.y, g[r].y);
          g[r].z = fmaf(gy, xb.z, g[r].z);
          g[r].w = fmaf(gy, xb.w, g[r].w);
        }
      }
      bool fast = true;
      if (!(a.dbg & 2))
#pragma unroll
      for (int r = 0; r < R; ++r) {
        fast &= adam_update_fast(p[r].x, m[r].x, v[r].x, g[r].x, coef);
        fast &= adam_update_fast(p[r].y, m[r].y, v[r].y, g[r].y, coef);
        fast &= adam_update_fast(p[r].z, m[r].z, v[r].z, g[r].z, coef);
        fast &= adam_update_fast(p[r].w, m[r].w, v[r].w, g[r].w, coef);
      }
      if (__any_sync(0xffffffffu, !fast)) {   // an operand outside the straight-line sequences' range: the library path
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int o = (warp * R + r) * LS_KC + lane * 4;
          p[r] = *reinterpret_cast<const float4*>(st + o);
          m[r] = *reinterpret_cast<const float4*>(st + LS_BOX + o);
          v[r] = *reinterpret_cast<const float4*>(st + 2 * LS_BOX + o);
          adam_update(p[r].x, m[r].x, v[r].x, g[r].x, coef);
          adam_update(p[r].y, m[r].y, v[r].y, g[r].y, coef);
          adam_update(p[r].z, m[r].z, v[r].z, g[r].z, coef);
          adam_update(p[r].w, m[r].w, v[r].w, g[r].w, coef);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (ok[r] && !(a.dbg & 1)) {
          const int row = n + warp * R + r;
          const long long off = model * a.sP + static_cast<long long>(row) * L.K + kk;
          // (streaming / evict-first stores for the moments, or for all three, measured: no change)
          *reinterpret_cast<float4*>(L.W + off) = p[r];
          *reinterpret_cast<float4*>(L.mW + off) = m[r];
          *reinterpret_cast<float4*>(L.vW + off) = v[r];
        }
      }
      __syncwarp();
      if (lane == 0) ls_bar_arrive(empty + pos.stage);
      pos.next(a.stages);
    }
    __syncwarp();
    if (lane == 0) ls_bar_arrive(xempty + j);
  }
#undef LS_ADAM_DECODE
  if (a.adv.st) {   // every CTA read the step state before its first update: the last one to finish may advance it
    asm volatile("bar.sync 1, %0;" ::"n"(LA_CONSUMERS) : "memory");
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(a.adv.counter, 1u) == gridDim.x - 1u) {
        *a.adv.counter = 0u;
        step_advance_apply(a.adv);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*LsEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static LsEncodeFn ls_encode_fn() {
  static LsEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<LsEncodeFn>(p);
  }
  return fn;
}

// fp32 [d2][d1][d0] (d0 contiguous, rows `ld1` elements apart, planes `ld2` elements apart), box [1][box1][128], no swizzle;
// out-of-range rows / columns are zero-filled
static int ls_make_map3(CUtensorMap* map, const float* ptr, int d0, int d1, int d2, long long ld1, long long ld2, int box1,
                        const char* who) {
  LsEncodeFn fn = ls_encode_fn();
  if (!fn) {
    set_error("%s: cuTensorMapEncodeTiled not available from the driver", who);
    return PGF_ERR_CUDA;
  }
  if (d2 <= 1 || ld2 <= 0) {
    d2 = 1;
    ld2 = static_cast<long long>(d1) * ld1;
  }
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld1) * 4u, static_cast<cuuint64_t>(ld2) * 4u};
  const cuuint32_t box[3] = {LS_KC, static_cast<cuuint32_t>(box1), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (%d) dims=%d,%d,%d strides=%lld,%lld", who, static_cast<int>(r), d0, d1, d2, ld1, ld2);
    return PGF_ERR_CUDA;
  }
  return PGF_OK;
}
// weights [n_models][N][K]: rows K apart, models `model_stride` elements apart, box of 32 rows
static int ls_make_map(CUtensorMap* map, const float* ptr, int N, int K, int n_models, long long model_stride, const char* who) {
  return ls_make_map3(map, ptr, K, N, n_models, K, model_stride, LS_ROWS, who);
}

// Rows per work item (a multiple of 32) so that the persistent grid is evenly loaded.  `units[i]` = (model, chunk, ...)
// combinations of layer i, each covering row_blocks[i] 32-row blocks; an item covers `blocks_per` of them.  Items of all
// layers have the same size, so contiguous item ranges are equal amounts of work.  Maximise items / (rounds * grid);
// among near-equals prefer the tallest items (fewer tile loads and reductions).
static int ls_pick_blocks(const long long* units, const int* row_blocks, int n_layers, int grid, int max_blocks, int item_cost = 2,
                          int split_cost = 4) {
  // estimated launch time in ring stages: rounds x (stages per item + per-item overhead), the overhead being the item's
  // epilogue (cross-warp reduction / tile hand-over, ~item_cost stages) plus, when a unit is split over several items, the
  // partial-sum exchange through L2 (~split_cost stages)
  int rb_max = 0;
  for (int i = 0; i < n_layers; ++i) rb_max = row_blocks[i] > rb_max ? row_blocks[i] : rb_max;
  if (max_blocks > rb_max) max_blocks = rb_max;
  long long best = -1;
  int best_bp = 1;
  for (int bp = max_blocks; bp >= 1; --bp) {
    long long items = 0;
    bool split = false;
    for (int i = 0; i < n_layers; ++i) {
      const int per = (row_blocks[i] + bp - 1) / bp;
      items += units[i] * per;
      split |= per > 1;
    }
    const long long rounds = (items + grid - 1) / grid;
    const long long cost = rounds * (bp + item_cost + (split ? split_cost : 0));
    if (best < 0 || cost < best) {
      best = cost;
      best_bp = bp;
    }
  }
  return best_bp;
}

// A layer's weights of ALL models of the launch fit L2 next to the streams passing through (<= 96 of 126 MB): keep them.
// Measured, 6 models per GPU (fc_layers.2 = 42 MB kept, fc_layers.0 = 127 MB streaming): 134.3 k -> 136.2 k model-samples/s;
// keeping both layers: 134.5 k; 12 models, both layers hinted: 143.4 k -> 145.7 k.
static int ls_w_keep(int n_models, int N, int K) {
  static const int env = getenv("PGF_LS_KEEP") ? atoi(getenv("PGF_LS_KEEP")) : -1;
  if (env >= 0) return env;
  return static_cast<double>(n_models) * N * K * 4.0 <= 96.0 * 1024 * 1024 ? 1 : 0;
}

static int ls_grid(int n_items) {
  const int g = num_sms();
  return n_items < g ? (n_items > 0 ? n_items : 1) : g;
}

static inline size_t ls_bar_bytes(int stages) { return (2 * static_cast<size_t>(stages) + 4) * sizeof(uint64_t); }

int linear_fwd(const LinFwdArgs& in, int n_models, cudaStream_t s) {
  LsFwdArgs a;
  a.bias = in.bias; a.sb = in.sb; a.Y = in.Y; a.ldy = in.ldy; a.sY = in.sY;
  a.B = in.B; a.N = in.N; a.K = in.K; a.act = in.act; a.n_models = n_models;
  a.row_blocks = (in.N + LS_ROWS - 1) / LS_ROWS;
  a.bchunks = (in.B + TB - 1) / TB;
  a.kchunks = (in.K + LS_KC - 1) / LS_KC;
  static const int dbg = getenv("PGF_LS_DBG") ? atoi(getenv("PGF_LS_DBG")) : 0;
  static const int stages_env = getenv("PGF_LS_STAGES") ? atoi(getenv("PGF_LS_STAGES")) : 0;
  a.dbg = dbg;
  a.stages = stages_env > 0 ? stages_env : 10;
  while (a.stages > 2 && static_cast<size_t>(a.stages) * LS_FSTAGE * 4 + ls_bar_bytes(a.stages) > LS_SMEM_MAX) --a.stages;
  const size_t smem = static_cast<size_t>(a.stages) * LS_FSTAGE * 4 + ls_bar_bytes(a.stages);
  a.x_models = (n_models > 1 && in.sX != 0) ? n_models : 1;
  a.w_keep = ls_w_keep(n_models, in.N, in.K);
  CUtensorMap tmW, tmX;
  int rc = ls_make_map(&tmW, in.W, in.N, in.K, n_models, in.sW, "pgf_linear_fwd");
  if (rc == PGF_OK) rc = ls_make_map3(&tmX, in.X, in.K, in.B, a.x_models, in.ldx, in.sX, TB, "pgf_linear_fwd");
  if (rc != PGF_OK) return rc;
  ensure_dynamic_smem(reinterpret_cast<const void*>(ls_fwd_kernel), smem);
  const int n_items = n_models * a.bchunks * a.row_blocks;
  launch(ls_fwd_kernel, dim3(ls_grid(n_items)), dim3(LS_THREADS), smem, s, tmW, tmX, a);
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_fwd");
  return PGF_OK;
}

// ---- dX
static void ls_dx_shape(int B, int N, int K, int n_models, LsDxArgs& a) {
  a.B = B; a.N = N; a.K = K; a.n_models = n_models;
  a.kchunks = (K + LS_KC - 1) / LS_KC;
  a.bchunks = (B + TB - 1) / TB;
  const int row_blocks = (N + LS_ROWS - 1) / LS_ROWS;
  const long long units = static_cast<long long>(n_models) * a.bchunks * a.kchunks;
  // the gradient tile [8][rows_per_split] must fit beside a >= 4-stage ring and the reduction buffer
  const int max_blocks = static_cast<int>((LS_SMEM_MAX - 4 * LS_BOX * 4 - LS_WARPS * TB * LS_KC * 4 - 256) / (2 * TB * LS_ROWS * 4));
  static const int bp_env = getenv("PGF_LS_DX_BP") ? atoi(getenv("PGF_LS_DX_BP")) : 0;
  int bp = ls_pick_blocks(&units, &row_blocks, 1, num_sms(), max_blocks);
  if (bp_env > 0) bp = bp_env < row_blocks ? bp_env : row_blocks;
  if (bp > max_blocks) bp = max_blocks;
  a.rows_per_split = bp * LS_ROWS;
  a.nsplit = (row_blocks + bp - 1) / bp;
  int stages = 8;
  auto need = [&](int st) {
    return static_cast<size_t>(st) * LS_BOX * 4 + static_cast<size_t>(LS_WARPS) * TB * LS_KC * 4 +
           2 * static_cast<size_t>(TB) * a.rows_per_split * 4 + ls_bar_bytes(st);
  };
  while (stages > 2 && need(stages) > LS_SMEM_MAX) --stages;
  a.stages = stages;
}

int linear_dx_counters(int B, int K, int n_models) {
  return n_models * ((B + TB - 1) / TB) * ((K + LS_KC - 1) / LS_KC);
}

size_t linear_dx_workspace(int B, int N, int K, int n_models) {
  LsDxArgs a;
  ls_dx_shape(B, N, K, n_models, a);
  const size_t units = static_cast<size_t>(n_models) * a.bchunks * a.kchunks;
  const size_t part = a.nsplit > 1 ? units * a.nsplit * TB * LS_KC * sizeof(float) : 0;
  return part + units * sizeof(unsigned int) + 256;   // + counters for callers that bring none
}

int linear_bwd_dx(const float* dY, long long ldy, long long sdY, const float* W, long long sW, const float* mask_src,
                  int mask_mode, long long ld_mask, long long s_mask, float* dX, long long ldx, long long sdX, int B, int N, int K,
                  int n_models, float* workspace, size_t workspace_bytes, cudaStream_t s, unsigned int* counters) {
  if (workspace_bytes < linear_dx_workspace(B, N, K, n_models)) {
    set_error("pgf_linear_bwd_dx: workspace too small");
    return PGF_ERR_WORKSPACE;
  }
  LsDxArgs a;
  ls_dx_shape(B, N, K, n_models, a);
  a.dY = dY; a.ldy = ldy; a.sdY = sdY; a.mask_src = mask_src; a.mask_mode = mask_mode; a.ld_mask = ld_mask; a.s_mask = s_mask;
  a.dX = dX; a.ldx = ldx; a.sdX = sdX;
  a.w_keep = ls_w_keep(n_models, N, K);
  const size_t units = static_cast<size_t>(n_models) * a.bchunks * a.kchunks;
  const size_t part = a.nsplit > 1 ? units * a.nsplit * TB * LS_KC * sizeof(float) : 0;
  a.partial = workspace;
  a.counters = counters;
  if (!counters && a.nsplit > 1) {   // ordinary entry point: counters live behind the partials and are zeroed per call
    a.counters = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(workspace) + ((part + 255) & ~static_cast<size_t>(255)));
    cudaMemsetAsync(a.counters, 0, units * sizeof(unsigned int), s);
  }
  CUtensorMap tm;
  const int rc = ls_make_map(&tm, W, N, K, n_models, sW, "pgf_linear_bwd_dx");
  if (rc != PGF_OK) return rc;
  const size_t smem = static_cast<size_t>(a.stages) * LS_BOX * 4 + static_cast<size_t>(LS_WARPS) * TB * LS_KC * 4 +
                      2 * static_cast<size_t>(TB) * a.rows_per_split * 4 + ls_bar_bytes(a.stages);
  if (smem > LS_SMEM_MAX) {
    set_error("pgf_linear_bwd_dx: N=%d too large for the shared-memory gradient tile", N);
    return PGF_ERR_UNSUPPORTED;
  }
  ensure_dynamic_smem(reinterpret_cast<const void*>(ls_dx_kernel), smem);
  const int n_items = static_cast<int>(units) * a.nsplit;
  launch(ls_dx_kernel, dim3(ls_grid(n_items)), dim3(LS_THREADS), smem, s, tm, a);
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_bwd_dx");
  return PGF_OK;
}

// ---- dW + Adam
int linear_adam_step(const LinAdamArgs& in, int n_models, cudaStream_t s) {
  if (in.B > TB) {
    set_error("pgf_linear_adam_step: the fused gradient+Adam kernel handles batches up to %d rows (got %d)", TB, in.B);
    return PGF_ERR_UNSUPPORTED;
  }
  LsAdamArgs a = {};
  LsAdamMaps tm;
  a.n_layers = in.n_layers; a.n_models = n_models; a.B = in.B; a.sP = in.sP; a.c = in.c; a.st = in.st; a.adv = in.adv;
  static const int dbg = getenv("PGF_LS_DBG") ? atoi(getenv("PGF_LS_DBG")) : 0;
  a.dbg = dbg;
  long long units[2] = {0, 0};
  int row_blocks[2] = {0, 0};
  for (int i = 0; i < in.n_layers; ++i) {
    const LinAdamLayer& l = in.l[i];
    if (l.K > 64 * LS_KC) {
      set_error("pgf_linear_adam_step: K <= %d", 64 * LS_KC);
      return PGF_ERR_ARG;
    }
    units[i] = static_cast<long long>(n_models) * ((l.K + LS_KC - 1) / LS_KC);
    row_blocks[i] = (l.N + LS_ROWS - 1) / LS_ROWS;
  }
  // one item height for every layer, so that contiguous item ranges are equal amounts of work
  static const int bp_env = getenv("PGF_LS_ADAM_BP") ? atoi(getenv("PGF_LS_ADAM_BP")) : 0;
  int bp = ls_pick_blocks(units, row_blocks, in.n_layers, num_sms(), 36, 1, 0);
  if (bp_env > 0) bp = bp_env;
  int n_items = 0, rows_max = 0;
  for (int i = 0; i < in.n_layers; ++i) {
    const LinAdamLayer& l = in.l[i];
    LsAdamLayerDev& d = a.l[i];
    d.dY = l.dY; d.ldy = l.ldy; d.sdY = l.sdY; d.X = l.X; d.ldx = l.ldx; d.sX = l.sX;
    d.W = l.W; d.mW = l.mW; d.vW = l.vW; d.bias = l.bias; d.mb = l.mb; d.vb = l.vb; d.N = l.N; d.K = l.K;
    d.kchunks = (l.K + LS_KC - 1) / LS_KC;
    d.w_keep = ls_w_keep(n_models, l.N, l.K);
    const int blocks_per = bp < row_blocks[i] ? bp : row_blocks[i];
    d.rsplit = (row_blocks[i] + blocks_per - 1) / blocks_per;
    d.rows_per_split = blocks_per * LS_ROWS;
    d.items = n_models * d.kchunks * d.rsplit;
    n_items += d.items;
    rows_max = d.rows_per_split > rows_max ? d.rows_per_split : rows_max;
    int rc = ls_make_map(&tm.w[i], l.W, l.N, l.K, n_models, in.sP, "pgf_linear_adam_step");
    if (rc == PGF_OK) rc = ls_make_map(&tm.m[i], l.mW, l.N, l.K, n_models, in.sP, "pgf_linear_adam_step");
    if (rc == PGF_OK) rc = ls_make_map(&tm.v[i], l.vW, l.N, l.K, n_models, in.sP, "pgf_linear_adam_step");
    if (rc != PGF_OK) return rc;
  }
  if (in.n_layers == 1) {
    tm.w[1] = tm.w[0]; tm.m[1] = tm.m[0]; tm.v[1] = tm.v[0];
  }
  a.rows_max = rows_max;
  int stages = 4;
  auto need = [&](int st) {
    return static_cast<size_t>(st) * 3 * LS_BOX * 4 + 2 * static_cast<size_t>(TB) * rows_max * 4 + 2 * TB * LS_KC * 4 + ls_bar_bytes(st) + 16;
  };
  while (stages > 2 && need(stages) > LS_SMEM_MAX) --stages;
  if (need(stages) > LS_SMEM_MAX) {
    set_error("pgf_linear_adam_step: layer too tall for the shared-memory gradient tile (%d rows per split)", rows_max);
    return PGF_ERR_UNSUPPORTED;
  }
  a.stages = stages;
  ensure_dynamic_smem(reinterpret_cast<const void*>(ls_adam_kernel), need(stages));
  launch(ls_adam_kernel, dim3(ls_grid(n_items)), dim3(LA_THREADS), need(stages), s, tm, a);
  PGF_CUDA_LAUNCH_CHECK("pgf_linear_adam_step");
  return PGF_OK;
}

}  // namespace pgf
