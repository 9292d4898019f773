// ns_block.cuh -- block-level building blocks shared by the exact arithmetic-coder kernel
// (ns_coder.cu) and the comparison codecs (ns_codecs.cu): 1024-thread CTAs, one row in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ns_coder.h"
#include "ns_math.cuh"

namespace {


typedef unsigned long long u64;

constexpr int NT = 1024;             // threads per CTA
constexpr int NWARPS = NT / 32;
constexpr int LIST_CAP = 512;        // resolve capacity of one selection bucket
constexpr int HIST_BYTES = 8192;     // 2048 x u32 (precision <= 31) or 1024 x u64
constexpr int SMEM_LIMIT = 232448;   // 227 KB opt-in maximum per CTA on sm_100

struct ListEntry {
  u64 pack;   // (orderable key << 32) | ~id : larger = earlier in the coder's order
  u64 w;      // weight (mass or count)
};

struct Scalars {
  u64 red[NWARPS];       // reduction scratch
  int list_count;
  int sel_bin;
  u64 sel_prefix;
  int res_idx;
  u64 res_before;
  u64 res_w;
  int res_found;
  u64 bar[8];            // mbarriers of the codecs' bulk row copy
};

constexpr int FIXED_BYTES = 26624;    // exp table, histogram, lists, scalars (both kernels); the row follows
static_assert(NS_EXP_N * 8 + HIST_BYTES + LIST_CAP * (int)sizeof(ListEntry) + 1024 <= FIXED_BYTES, "exact smem layout");
static_assert(sizeof(Scalars) <= 1024, "scalar block too large");
constexpr int MAX_VOCAB = (SMEM_LIMIT - FIXED_BYTES) / 4 - 8;

__constant__ double c_exp_tab[NS_EXP_N] = {NS_EXP_TAB_VALUES};
// the same table in global memory: a lookup with a different index in every lane is one L1 access there, but up to 32
// serialised passes through the constant cache
__device__ const double g_exp_tab[NS_EXP_N] = {NS_EXP_TAB_VALUES};

// ------------------------------------------------------------------------------------
// block-wide reductions, bit-deterministic: xor-butterfly inside the warp (commutative,
// so every lane holds the same bits), then warp partials combined in index order.
// ------------------------------------------------------------------------------------
struct OpAddD { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpAddU { __device__ u64 operator()(u64 a, u64 b) const { return a + b; } };
struct OpMaxU { __device__ u64 operator()(u64 a, u64 b) const { return a > b ? a : b; } };
struct OpMinU { __device__ u64 operator()(u64 a, u64 b) const { return a < b ? a : b; } };

template <class Op>
__device__ __forceinline__ u64 block_reduce_u(u64 v, Op op, u64* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  // NWARPS == 32 partials: one per lane, a second butterfly (the operators are exact and commutative, so the
  // order does not matter and every thread ends with the same value)
  u64 r = scratch[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, o));
  return r;
}
static_assert(NWARPS == 32, "block_reduce_u combines one partial per lane");

__device__ __forceinline__ double block_sum_d(double v, u64* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = (u64)__double_as_longlong(v);
  __syncthreads();
  double r = __longlong_as_double((long long)scratch[0]);
#pragma unroll 1
  for (int w = 1; w < NWARPS; ++w) r = r + __longlong_as_double((long long)scratch[w]);
  return r;
}

__device__ __forceinline__ u64 pack_of(float key, int id) {
  return ((u64)ns_f32_orderable(key) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)id);
}
__device__ __forceinline__ float key_of_pack(u64 p) {
  uint32_t u = (uint32_t)(p >> 32);
  uint32_t bits = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  return __uint_as_float(bits);
}
__device__ __forceinline__ int id_of_pack(u64 p) { return (int)(0xFFFFFFFFu - (uint32_t)p); }

__device__ __forceinline__ int bin_of(float key, float m, float scale, int nb) {
  float d = (m - key) * scale;          // monotone non-increasing in key
  d = fminf(d, (float)(nb - 1));
  return (int)d;
}

// ------------------------------------------------------------------------------------
// histogram selection: first position in the coder's order whose inclusive cumulative
// weight exceeds tau.  The histogram has been filled by the caller.
// ------------------------------------------------------------------------------------
template <typename HistT, int NB>
__device__ void sel_locate(const HistT* hist, u64 tau, Scalars* sc) {
  constexpr int BPT = NB / NT;          // bins per thread (2 for u32, 1 for u64)
  static_assert(BPT >= 1, "histogram smaller than the CTA");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64 local[BPT];
  u64 tsum = 0;
#pragma unroll
  for (int b = 0; b < BPT; ++b) { local[b] = (u64)hist[tid * BPT + b]; tsum += local[b]; }
  // inclusive scan over threads
  u64 inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u64 t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) sc->red[warp] = inc;
  if (tid == 0) { sc->sel_bin = -1; sc->sel_prefix = 0; }
  __syncthreads();
  u64 woff = 0;
#pragma unroll 1
  for (int w = 0; w < warp; ++w) woff += sc->red[w];
  u64 excl = woff + inc - tsum;
#pragma unroll
  for (int b = 0; b < BPT; ++b) {
    if (local[b] != 0 && excl <= tau && tau < excl + local[b]) {
      sc->sel_bin = tid * BPT + b;
      sc->sel_prefix = excl;
    }
    excl += local[b];
  }
  __syncthreads();
}

// Exact resolution among the collected entries of the target bucket.  The n x n comparisons are spread over the
// whole CTA: a group of `ways` adjacent lanes shares one entry, each lane scans every ways-th other entry, and
// the partial weights meet by shuffles (integer sums: exact in any order).
__device__ void sel_resolve(const ListEntry* list, int n, u64 tau, u64 prefix, Scalars* sc) {
  if (threadIdx.x == 0) sc->res_found = 0;
  __syncthreads();
  int ways = 1;
  while (ways < 8 && n * ways * 2 <= NT) ways *= 2;          // 1, 2, 4 or 8 lanes per entry (uniform across the CTA)
  const int sub = threadIdx.x & (ways - 1);
  for (int base = 0; base < n; base += NT / ways) {          // one trip unless n > NT (never: n <= LIST_CAP)
    const int c = base + threadIdx.x / ways;
    const bool live = c < n;
    const u64 pc = live ? list[c].pack : 0ull;
    u64 before = 0;
    if (live) {
      for (int o = sub; o < n; o += ways) {
        const u64 po = list[o].pack;
        if (po > pc) before += list[o].w;
      }
    }
    for (int o = 1; o < ways; o <<= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (live && sub == 0) {
      const u64 wc = list[c].w;
      before += prefix;
      if (wc != 0 && before <= tau && tau < before + wc) {
        sc->res_idx = id_of_pack(pc);
        sc->res_before = before;
        sc->res_w = wc;
        sc->res_found = 1;
      }
    }
  }
  __syncthreads();
}


}  // namespace
