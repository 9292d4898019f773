// ns_math.cuh -- scalar building blocks of the coder step, usable on host and device
// (the host build is exercised by tests/test_native_math.py through csrc/ns_hosttest.cpp).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define NS_HD __host__ __device__ __forceinline__
#else
#define NS_HD inline
#include <cmath>
#endif

#include "ns_exp_table.inc"

#define NS_EXP_L 9
#define NS_EXP_N 512
#define NS_EXP_UNDERFLOW (-708.0)

NS_HD double ns_u64_as_double(uint64_t u) {
  double d;
#if defined(__CUDA_ARCH__)
  d = __longlong_as_double((long long)u);
#else
  memcpy(&d, &u, 8);
#endif
  return d;
}
NS_HD uint64_t ns_double_as_u64(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
#endif
}
NS_HD double ns_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}

// exp(a) for -708 <= a <= 0 (caller guarantees the range), fp64, about 1 ulp.
//   n = rint(a * 512/ln2), r = a - n*ln2/512 (two-term Cody-Waite, n*HI exact),
//   exp(a) = 2^(n>>9) * T[n&511] * (1 + r + r^2/2 + r^3/6 + r^4/24),  |r| <= 6.8e-4
// `tab` points at NS_EXP_TAB (shared memory on the device).  Branch-free: 10 fp64 ops.
NS_HD double ns_exp64_core(double a, const double* __restrict__ tab) {
  const double magic = 6755399441055744.0;           // 1.5 * 2^52
  double t = ns_fma(a, NS_512_OVER_LN2, magic);
  int32_t n = (int32_t)(uint32_t)ns_double_as_u64(t);  // low word = rint(a*512/ln2), two's complement
  double nd = t - magic;
  double r = ns_fma(nd, -NS_LN2_512_HI, a);
  r = ns_fma(nd, -NS_LN2_512_LO, r);
  double T = tab[n & (NS_EXP_N - 1)];
  double q = ns_fma(r, 1.0 / 24.0, 1.0 / 6.0);
  q = ns_fma(q, r, 0.5);
  double r2 = r * r;
  double p = ns_fma(q, r2, r);
  double e = ns_fma(T, p, T);
  // scale by 2^k through the exponent field (result stays normal for a >= -708)
#if defined(__CUDA_ARCH__)
  const int hi = __double2hiint(e) + ((n & ~(NS_EXP_N - 1)) << (20 - NS_EXP_L));   // high word += k << 20, k = n >> 9
  return __hiloint2double(hi, __double2loint(e));
#else
  const int32_t k = n >> NS_EXP_L;                   // arithmetic shift = floor
  uint64_t bits = ns_double_as_u64(e) + ((uint64_t)(int64_t)k << 52);
  return ns_u64_as_double(bits);
#endif
}

// exp(a) for any a <= 0: exactly 0 below -708 (also for the -inf mask and NaN).
NS_HD double ns_exp64_neg(double a, const double* __restrict__ tab) {
  if (!(a >= NS_EXP_UNDERFLOW)) return 0.0;
  return ns_exp64_core(a, tab);
}

// Order-preserving map fp32 -> u32 (ascending).
NS_HD uint32_t ns_f32_orderable(float f) {
  uint32_t u;
#if defined(__CUDA_ARCH__)
  u = __float_as_uint(f);
#else
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// "x sorts before y" in the coder's order: larger logit first, equal logits by lower id
// (reference: torch.sort descending, code_base/arithmetic.py:127; tie order is ours).
NS_HD bool ns_before(float kx, int32_t ix, float ky, int32_t iy) {
  return (kx > ky) || (kx == ky && ix < iy);
}

// Shared-prefix count and interval rescale (code_base/arithmetic.py:179-190 with the
// num_same_from_beg cap of code_base/utils.py:59-64: never more than precision-1 bits).
//   nb = new_bottom, nt = new_top (exclusive).  Returns n; writes the new [lo, hi).
NS_HD int ns_interval_update(uint64_t nb, uint64_t nt, int precision, uint64_t* lo, uint64_t* hi) {
  uint64_t tm1 = nt - 1;
  uint64_t x = nb ^ tm1;                             // differing bits among the low `precision`
  int n;
  if (x == 0) {
    n = precision - 1;
  } else {
#if defined(__CUDA_ARCH__)
    int msb = 63 - __clzll((long long)x);
#else
    int msb = 63 - __builtin_clzll(x);
#endif
    n = precision - 1 - msb;                         // bits above the first difference
    if (n > precision - 1) n = precision - 1;
    if (n < 0) n = 0;
  }
  uint64_t mask = (precision >= 64) ? ~0ull : ((1ull << precision) - 1ull);
  *lo = (nb << n) & mask;
  *hi = (((tm1 << n) & mask) | ((1ull << n) - 1ull)) + 1ull;
  return n;
}

// `count` (<= 48) bits of an MSB-first packed bit string starting at bit `pos`,
// zero-padded past `len` (code_base/arithmetic.py:168-171).  Returned MSB-first as an integer.
NS_HD uint64_t ns_read_bits(const uint32_t* words, int32_t pos, int32_t len, int count) {
  uint64_t v = 0;
  int32_t avail = len - pos;
  if (avail < 0) avail = 0;
  int take = count < avail ? count : (int)avail;
  int got = 0;
  while (got < take) {
    int32_t b = pos + got;
    uint32_t w = words[b >> 5];
    int off = b & 31;
    int chunk = 32 - off;
    if (chunk > take - got) chunk = take - got;
    uint32_t part = (w << off) >> (32 - chunk);
    v = (v << chunk) | part;
    got += chunk;
  }
  return v << (count - take);
}

// Append the `count` (<= 48) low bits of `value`, MSB first, at bit position `pos`.
NS_HD void ns_write_bits(uint32_t* words, int32_t pos, uint64_t value, int count) {
  int done = 0;
  while (done < count) {
    int32_t b = pos + done;
    int off = b & 31;
    int chunk = 32 - off;
    if (chunk > count - done) chunk = count - done;
    uint32_t part = (uint32_t)((value >> (count - done - chunk)) & ((chunk == 32) ? 0xFFFFFFFFu : ((1u << chunk) - 1u)));
    uint32_t shift = 32 - off - chunk;
    uint32_t m = ((chunk == 32) ? 0xFFFFFFFFu : ((1u << chunk) - 1u)) << shift;
    uint32_t w = words[b >> 5];
    words[b >> 5] = (w & ~m) | (part << shift);
    done += chunk;
  }
}
