// ns_fast.cuh -- single-exp-pass arithmetic-coder step (the throughput path), sm_100a.
// Included by ns_coder.cu after the shared definitions (u64, pack_of, finish_*).
//
// Persistent CTAs (one per SM, 512 threads), each looping over rows.  Per row (threshold form of the cutoff):
//   L   the row is pulled into shared memory by the bulk-copy engine (cp.async.bulk, one copy, one
//       mbarrier) -- issued by the previous row as soon as that row has read the buffer for the last
//       time; then an fp32 online softmax estimate (row max, its lowest id, sum of exp) runs over it.  The row after this one is prefetched into L2 (bulk prefetch).
//   P1  ONE fp64 exp per element (10 fp64 ops): exact sum of all e_i in a fixed order, exact sum of
//       the provisionally-cut ones, elements within 2^-10 of the provisional cutoff go to a small
//       list with their exact e ; the word is overwritten in place by the high word of e_i (0 if not kept)
//   FIX exact normaliser -> the provisional cutoff is verified, list elements classified exactly,
//       S_kept and C = range / S_kept exact
//   P2  q_i = rint(e_i * C) from the truncated e_i with a rigorous interval test (2 fp64 FMAs);
//       the few undecidable ones are redone exactly from the original logit (L2 hit) ;
//       encode: integer mass histogram over 2048 monotone buckets of the fp32 bit pattern;
//       decode: no histogram -- the mass ranked before the observed token is a conditional sum
//   SEL (encode) bucket prefix -> gather the target bucket -> exact order (e32, original logit, id)
//   UPD shared-prefix bits + interval rescale (finish_encode / finish_decode)
// Rank form (top-k binds, 2 <= topk <= 512; fast_rank_row, RANK instantiation): after L, no fp64 pass over the
// row -- sampled histogram -> bound on the top-k key -> candidate list -> exact selection of the topk keys ->
// exp / widths / prefix sums on topk elements.
// Anything unusual (list overflows, estimate outside its guard band, rank form beyond the lists) queues
// the row in slow_ws; the exact multi-pass kernel then redoes it with the same integers.

// One CTA per SM, 512 threads, the row resident in shared memory (bulk copy).
#ifndef NSF_FT
#define NSF_FT 512
#endif
#ifndef NSF_MIN_CTAS
#define NSF_MIN_CTAS 1
#endif
constexpr int FT = NSF_FT;           // threads per CTA
constexpr int FW = FT / 32;
constexpr int F_NB = 2048;           // histogram buckets (u32 masses: precision <= 31)
constexpr int F_BPT = F_NB / FT;     // buckets per thread in the scan
constexpr int F_BAND_CAP = 128;
constexpr int F_U_CAP = 256;
constexpr int F_C_CAP = 256;
constexpr int F_PIECES = 1;          // bulk-copy pieces per row (F_SUB * 3 * FT float4 each)
constexpr int F_SUB = 9;             // blocks of 3 float4 per thread in a piece
constexpr int F_MIN_VOCAB = 256;     // below this the exact kernel is used
constexpr float F_BAND_EPS = 0.0009765625f;   // 2^-10 half-width (in log units) of the exact-list band
constexpr uint32_t F_TOP = 0xFF000000u;       // packed e of the row maximum (e == 1.0)

// hand-over reasons (status bits 8..15, diagnostics only)
enum { F_WHY_EST = 1, F_WHY_BAND = 2, F_WHY_VERIFY = 3, F_WHY_RANK = 4, F_WHY_ULIST = 5, F_WHY_BUCKET = 6 };

struct BandEntry { int id; int kept; double e; };
struct CandEntry { uint32_t ebits; int id; uint32_t w; float key; };   // 16 B

// per-row scalars, fetched one row ahead by a helper lane so no global latency sits on the row's path
struct RowMeta {
  u64 lo, hi, window;
  int slot, cursor, mlen, tok;
  int phase, olen;                   // decode: bits recovered so far
  uint32_t oword, pad;               // decode: the partly filled output word at olen >> 5
};

struct FScal {
  RowMeta meta[2];
  u64 red[3 * FW];
  u64 bar[F_PIECES];                 // mbarriers of the row copy
  float sum32; int remax; float M; int top_id;
  int band_n; int u_n; int c_n; int bail;
  int issued_row; int pad_issue;     // row whose bulk copy is already in flight (issued before the previous row ended)
  u64 band_cut_int;
  int band_kept_n; int sh;
  int sel_bin; u64 sel_prefix;
  int res_idx; u64 res_before; u64 res_w; int res_found;
  float kappa_lo, kappa_hi, clamp_key; int band_E;
};
static_assert(sizeof(FScal) <= 2048, "FScal too large");
static_assert(NS_EXP_N * 8 + F_NB * 4 + F_BAND_CAP * 16 + F_U_CAP * 4 + 2 * F_C_CAP * 16 + 2048 <= FIXED_BYTES, "fast smem layout");

__device__ __forceinline__ float f_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void f_prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}
__device__ __forceinline__ uint32_t f_smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void f_mbar_init(u64* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(f_smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void f_mbar_expect_tx(u64* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(f_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_bulk_g2s(void* dst, const void* src, uint32_t bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(f_smem_addr(dst)), "l"(src), "r"(bytes), "r"(f_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void f_mbar_wait(u64* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" :: "r"(f_smem_addr(bar)), "r"(parity) : "memory");
}

// Bulk copy of row `row` into the shared-memory row (interior float4 chunks 1 .. W4-2, F_PIECES pieces, one
// mbarrier each).  One thread.  Every generic-proxy access of the buffer's previous content must be behind a
// CTA barrier; the proxy fence orders them before the async-proxy writes.
__device__ __forceinline__ void f_issue_row(const ns_ac_params& P, int row, u64* bar, float* words, int* issued_row);

// The exp pass leaves, in place of each kept logit, a 32-bit truncation of its fp64 e: the double's
// bits 59..28 (low 8 exponent bits + 24 mantissa bits; for 2^-255 < e <= 1 the four bits above are the
// constant 0011).  It is monotone in e, accurate to 2^-24 (toward zero), and turns back into a double
// with two shifts -- no trip through the conversion unit.  Carried in the float4 row as a bit pattern;
// 0 = not kept (unpacks to 2^-255: its bin width rounds to 0).
__device__ __forceinline__ float f_pack_e(double e) {
  return __uint_as_float(__funnelshift_l((uint32_t)__double2loint(e), (uint32_t)__double2hiint(e), 4));
}
__device__ __forceinline__ double f_unpack_e(float w) {
  const uint32_t b = __float_as_uint(w);
  return __hiloint2double((int)__funnelshift_r(b, 0x3u, 4), (int)(b << 28));
}

// hist[bin] += q, unconditionally: q == 0 (not kept) adds nothing, and without a predicate the
// compiler batches the eight address computations and reductions of an iteration (no branches).
__device__ __forceinline__ void f_hist_add(uint32_t* hist, uint32_t bin, uint32_t q) {
  // not-kept elements all map to one bucket index: spread their (zero) adds over the lanes' own buckets
  atomicAdd(hist + (q ? bin : (threadIdx.x & (F_NB - 1))), q);
}

// queue a row for the exact kernel: slow_ws = {count, done, rows...}
__device__ __forceinline__ void hand_over(const ns_ac_params& P, int32_t* slow_ws, int row, int why) {
  const int s = atomicAdd(&slow_ws[0], 1);
  slow_ws[2 + s] = row;
  if (P.status) atomicOr(&P.status[row], NS_ST_EST_RETRY | (why << 8));   // informational
}

// max of a and min of b in one pass (one pair of barriers)
__device__ __forceinline__ void f_reduce_maxmin(u64& a, u64& b, u64* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const u64 x = __shfl_xor_sync(0xffffffffu, a, o), y = __shfl_xor_sync(0xffffffffu, b, o);
    a = x > a ? x : a;
    b = y < b ? y : b;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { scratch[threadIdx.x >> 5] = a; scratch[FW + (threadIdx.x >> 5)] = b; }
  __syncthreads();
  u64 ra = scratch[0], rb = scratch[FW];
#pragma unroll
  for (int w = 1; w < FW; ++w) {
    const u64 x = scratch[w], y = scratch[FW + w];
    ra = x > ra ? x : ra;
    rb = y < rb ? y : rb;
  }
  a = ra; b = rb;
}

template <class Op>
__device__ __forceinline__ u64 f_reduce_u(u64 v, Op op, u64* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  u64 r = scratch[0];
#pragma unroll
  for (int w = 1; w < FW; ++w) r = op(r, scratch[w]);
  return r;
}
__device__ __forceinline__ float f_sum_f(float v, u64* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = (u64)__float_as_uint(v);
  __syncthreads();
  float r = __uint_as_float((uint32_t)scratch[0]);
#pragma unroll
  for (int w = 1; w < FW; ++w) r = r + __uint_as_float((uint32_t)scratch[w]);
  return r;
}
// two fp64 sums and one integer sum with a single pair of barriers; fixed order -> deterministic bits
__device__ __forceinline__ void f_sum_ddu(double& a, double& b, u64& c, u64* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = a + __shfl_xor_sync(0xffffffffu, a, o);
    b = b + __shfl_xor_sync(0xffffffffu, b, o);
    c = c + __shfl_xor_sync(0xffffffffu, c, o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    scratch[w] = (u64)__double_as_longlong(a);
    scratch[FW + w] = (u64)__double_as_longlong(b);
    scratch[2 * FW + w] = c;
  }
  __syncthreads();
  double ra = __longlong_as_double((long long)scratch[0]);
  double rb = __longlong_as_double((long long)scratch[FW]);
  u64 rc = scratch[2 * FW];
#pragma unroll
  for (int w = 1; w < FW; ++w) {
    ra = ra + __longlong_as_double((long long)scratch[w]);
    rb = rb + __longlong_as_double((long long)scratch[FW + w]);
    rc = rc + scratch[2 * FW + w];
  }
  a = ra; b = rb; c = rc;
}

__device__ __forceinline__ void f_issue_row(const ns_ac_params& P, int row, u64* bar, float* words, int* issued_row) {
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  const int NI = ((mis + P.V + 3) >> 2) - 2;
  constexpr int PC = F_SUB * 3 * NSF_FT;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const char* src = reinterpret_cast<const char*>(g - mis) + 16;
  char* dst = reinterpret_cast<char*>(words) + 16;
  for (int k = 0; k < F_PIECES; ++k) {
    const int c0 = k * PC;
    int n = NI - c0;
    if (n > PC) n = PC;
    if (n > 0) {
      f_mbar_expect_tx(&bar[k], (uint32_t)n * 16u);
      f_bulk_g2s(dst + (size_t)c0 * 16, src + (size_t)c0 * 16, (uint32_t)n * 16u, &bar[k]);
    }
  }
  *issued_row = row;
}

struct FastSmem {
  double* tab; uint32_t* hist; BandEntry* band; int* ulist; CandEntry* clist; FScal* sc; float* words;
};

__device__ __forceinline__ RowMeta f_load_meta(const ns_ac_params& P, int row, int mode) {
  RowMeta m;
  m.phase = P.phase ? (int)P.phase[row] : NS_PHASE_CODING;
  m.slot = P.ntok ? P.ntok[row] : 0;
  m.lo = P.lo[row]; m.hi = P.hi[row];
  m.cursor = 0; m.mlen = 0; m.window = 0; m.tok = -1; m.pad = 0; m.olen = 0; m.oword = 0;
  if (mode == MODE_ENC) {
    m.cursor = P.cursor[row];
    m.mlen = P.msg_len[row];
    m.window = ns_read_bits(P.msg + (size_t)row * P.msg_stride, m.cursor, m.mlen, P.precision);   // :168-171
  } else {
    const int total = P.ntok_total ? P.ntok_total[row] : 0x7fffffff;
    m.mlen = total;
    if (m.slot < total) m.tok = P.token_in[(size_t)row * P.token_stride + m.slot];
    m.olen = P.out_len[row];
    m.oword = P.out_bits[(size_t)row * P.out_stride + (m.olen >> 5)];
  }
  return m;
}

// finish_decode (ns_coder.cu) with the stream's scalars and its partly filled output word already in registers:
// stores only, nothing on the row's path waits for global memory.  Needs ntok_total (else the caller uses
// finish_decode).  Words past the partial one are fresh (the output buffer is append-only and zero-initialised).
__device__ __forceinline__ void f_finish_decode(const ns_ac_params& P, int row, int slot, bool in_range, u64 nb, u64 nt,
                                                u64 k0, u64 Q, int total, int olen, uint32_t oword) {
  uint64_t nlo, nhi;
  const int n = ns_interval_update(nb, nt, P.precision, &nlo, &nhi);
  P.lo[row] = nlo; P.hi[row] = nhi;
  const bool last = slot == total - 1;
  if (P.ntok) P.ntok[row] = slot + 1;
  if (P.phase && slot + 1 >= total) P.phase[row] = NS_PHASE_DONE;
  const int count = last ? P.precision : n;                  // :356-359
  const u64 value = last ? nb : (n > 0 ? (nt - 1) >> (P.precision - n) : 0ull);
  uint32_t* ob = P.out_bits + (size_t)row * P.out_stride;
  int done = 0;
  uint32_t w = oword;
  while (done < count) {
    const int b = olen + done, off = b & 31;
    int chunk = 32 - off;
    if (chunk > count - done) chunk = count - done;
    const uint32_t mask = chunk == 32 ? 0xFFFFFFFFu : ((1u << chunk) - 1u);
    const uint32_t part = (uint32_t)(value >> (count - done - chunk)) & mask;
    ob[b >> 5] = w | (part << (32 - off - chunk));
    w = 0;
    done += chunk;
  }
  P.out_len[row] = olen + count;
  if (P.nbits_out) P.nbits_out[row] = (uint8_t)n;
  if (!in_range && P.status) atomicOr(&P.status[row], NS_ST_OUT_OF_RANGE);
  if (P.trace) { uint64_t* t = P.trace + (size_t)row * 4; t[0] = nb; t[1] = nt; t[2] = k0; t[3] = Q; }
}

// phase timers (thread 0 only, active when P.prof is given)
struct PhaseClock {
  bool on; long long last; u64 acc[16];
  __device__ __forceinline__ void start() { if (on) last = clock64(); }
  __device__ __forceinline__ void mark(int k) { if (on) { const long long t = clock64(); acc[k] += (u64)(t - last); last = t; } }
};


// ------------------------------------------------------------------------------------------------
// Rank form of the cutoff (code_base/arithmetic.py:75 with top-k binding: k = topk because more than
// topk tokens have p >= 1/range).  Only the topk largest logits matter, so no pass over the row touches
// fp64: a count histogram of the keys that are certainly above the cutoff, the bucket of position
// topk-1, a gather of the buckets before it (grouped by bucket, so ordering is local) plus an exact
// resolution of the boundary bucket, then exp / bin widths / prefix sums for topk elements only.
// Returns 0 when the row is not certainly in rank form (the caller goes on with the threshold form,
// the histogram is clean again), 1 when the row is finished or handed over.
// ------------------------------------------------------------------------------------------------
constexpr int F_K_CAP = 512;         // topk the path holds (the candidate list area, one thread per kept token)
constexpr int F_RB_CAP = 128;        // entries of the boundary bucket (the band list area)
static_assert(F_K_CAP <= FT && F_K_CAP <= 2 * F_C_CAP && F_RB_CAP <= F_BAND_CAP, "rank-form capacities");
static_assert(F_K_CAP * 16 <= F_NB * 4, "sorted arrays live in the histogram area");

template <bool UNIT_TEMP, int MODE>
__device__ __noinline__ int fast_rank_row(const ns_ac_params& P, int32_t* slow_ws, const int row, const int nrow,
                                          const u64 m_lo, const u64 m_hi, const u64 m_window, const int m_slot,
                                          const int m_cursor, const int m_mlen, const int m_tok,
                                          uint32_t* hist, FScal* sc, CandEntry* top, BandEntry* bnd, const float* words,
                                          const double* tab, const int mis, const int W4,
                                          const float M, const int top_id, const float kappa_r,
                                          const float clamp_key, const double dm) {
  // the pointers come through a real call: tell the compiler they are shared-memory addresses (LDS / ATOMS,
  // not generic loads and atomics)
  __builtin_assume(__isShared(hist)); __builtin_assume(__isShared(sc)); __builtin_assume(__isShared(top));
  __builtin_assume(__isShared(bnd)); __builtin_assume(__isShared(words)); __builtin_assume(__isShared(tab));
  const float4* w4 = reinterpret_cast<const float4*>(words);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, K = P.topk;
  long long tq = clock64();                                // phase timers (slots 11..14, only with P.prof)
  auto lap = [&](int k) { if (P.prof && tid == 0) { const long long t = clock64(); atomicAdd((unsigned long long*)&P.prof[k], (unsigned long long)(t - tq)); tq = t; } };
  const float span = M - kappa_r;
  if (!(span > 0.0f)) return 0;
  // monotone bucket of a key without a conversion instruction: (M - v) * scale, clamped, rounded by the 2^23
  // trick; the low mantissa bits are the bucket
  const float b_max = 8388608.0f + (float)(F_NB - 1);
  auto bucket_of = [&](float v, float scale, float off) -> int {
    return __float_as_int(fmaxf(fminf(fmaf(-v, scale, off), b_max), 8388608.0f)) & (F_NB - 1);
  };
  // block scan of the histogram: F_BPT buckets per thread; returns this thread's exclusive prefix and the total
  uint32_t hloc[F_BPT];
  auto scan_hist = [&](uint32_t* excl_out, uint32_t* total_out) {
    uint32_t tsum = 0;
#pragma unroll
    for (int b = 0; b < F_BPT; ++b) { hloc[b] = hist[tid * F_BPT + b]; tsum += hloc[b]; }
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();                                       // the previous user of sc->red is done
    if (lane == 31) sc->red[warp] = inc;
    if (tid == 0) { sc->sel_bin = -1; sc->sel_prefix = 0; }
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < FW; ++w) { const uint32_t x = (uint32_t)sc->red[w]; if (w < warp) woff += x; total += x; }
    *excl_out = woff + inc - tsum;
    *total_out = total;
  };
  // bucket holding position `pos` of the scanned histogram -> sc->sel_bin / sel_prefix (after a barrier)
  auto locate_pos = [&](uint32_t excl, uint32_t pos) {
#pragma unroll
    for (int b = 0; b < F_BPT; ++b) {
      if (hloc[b] != 0 && excl <= pos && pos < excl + hloc[b]) { sc->sel_bin = tid * F_BPT + b; sc->sel_prefix = excl; }
      excl += hloc[b];
    }
    __syncthreads();
  };
  auto clear_hist = [&]() {
#pragma unroll
    for (int b = 0; b < F_BPT; ++b) hist[tid * F_BPT + b] = 0;
  };

  // ---- level 1: count histogram of a 1/8 sample (whole warp iterations, spread over the ids) of the keys
  // certainly above the cutoff -> a key k_c that bounds the top-k from below with ~1.5 K candidates above it.
  // Branch-free inside a sampled iteration: keys below the cutoff add 0 to a bucket of the lane's own.
  const float rscale = (float)F_NB / span, b_off = M * rscale + 8388608.0f;
  {
    int it = 0;
    for (int cw = tid - lane; cw < W4; cw += 2 * FT, ++it) {
      if (((it + warp) & 7) != 0) continue;                // warp-uniform
      const int c = cw + lane;
      const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      const float4 v = (c < W4) ? w4[c] : ninf;
      const float4 w = (c + FT < W4) ? w4[c + FT] : ninf;
      const float xs[8] = {v.x, v.y, v.z, v.w, w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t inc = xs[e] >= kappa_r ? 1u : 0u;
        atomicAdd(hist + (inc ? bucket_of(xs[e], rscale, b_off) : (tid & (F_NB - 1))), inc);
      }
    }
  }
  __syncthreads();
  lap(11);
  float k_c = kappa_r;
  {
    uint32_t excl, total_s;
    scan_hist(&excl, &total_s);
    const uint32_t r_s = (uint32_t)((3 * K) / 16 + 12);    // 1.5 K / 8 plus a margin of > 5 sigma of the sample count
    if (total_s > r_s) {
      locate_pos(excl, r_s - 1u);
      // lower key edge of that sample bucket (round-to-nearest buckets: b covers [b - 0.5, b + 0.5)), one more
      // bucket of slack for the rounding of the offset
      k_c = fmaxf(kappa_r, M - ((float)sc->sel_bin + 1.5f) / rscale);
    }
    __syncthreads();
    clear_hist();
  }
  // ---- list the ids of the keys >= k_c.  Hits are rare per element but present in almost every warp iteration:
  // a branch-free sweep marks, per thread, the iterations whose eight keys hold a hit (one bit each); then every
  // thread revisits its few marked iterations.  The list lives where the gathered entries go later (each thread
  // has its listed tokens in registers by then).
  int* cand = reinterpret_cast<int*>(top);
  constexpr int CAND_CAP = 4 * FT;                         // <= 2 * F_C_CAP * 16 / 4 ints
  static_assert(CAND_CAP * 4 <= 2 * F_C_CAP * 16, "candidate list must fit the gathered-entry area");
  uint32_t marks = 0;
  {
    int it = 0;
    for (int c = tid; c < W4; c += 2 * FT, ++it) {
      const float4 v = w4[c];
      const float4 w = (c + FT < W4) ? w4[c + FT] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      const float mx = fmaxf(fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)), fmaxf(fmaxf(w.x, w.y), fmaxf(w.z, w.w)));
      marks |= (mx >= k_c ? 1u : 0u) << it;
    }
  }
  while (marks) {
    const int it = __ffs(marks) - 1;
    marks &= marks - 1u;
    const int c = tid + it * 2 * FT;
    const float4 v = w4[c];
    const float4 w = (c + FT < W4) ? w4[c + FT] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    const float xs[8] = {v.x, v.y, v.z, v.w, w.x, w.y, w.z, w.w};
    uint32_t hits = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) hits |= (xs[e] >= k_c ? 1u : 0u) << e;
    while (hits) {                                         // one or two per revisit
      const int e = __ffs(hits) - 1;
      hits &= hits - 1u;
      const int s2 = atomicAdd(&sc->c_n, 1);
      if (s2 < CAND_CAP) cand[s2] = 4 * (e < 4 ? c : c + FT) - mis + (e & 3);
    }
  }
  __syncthreads();
  const int ncand = sc->c_n;
  if (ncand > CAND_CAP || ncand <= K) {
    // too many ties / a flat row for the list, or (k_c == kappa_r) not more than K keys above the cutoff: not a
    // row for this path.  The histogram is clean; the list counter goes back to zero for the threshold form.
    __syncthreads();
    if (tid == 0) sc->c_n = 0;
    if (ncand > CAND_CAP) { if (tid == 0) hand_over(P, slow_ws, row, F_WHY_BUCKET); return 1; }
    return 0;
  }
  lap(12);
  // ---- level 2: count histogram of the listed keys over [k_c, M], bucket of position K-1, gather: buckets before
  // it land grouped by bucket at their prefix (count | exclusive prefix << 16 per bucket), the boundary bucket
  // goes to its own list
  const float span2 = M - k_c;
  const float rscale2 = span2 > 0.0f ? (float)F_NB / span2 : 0.0f, b_off2 = M * rscale2 + 8388608.0f;
  float ck[4]; int cid[4], cb[4];                          // this thread's (up to four) listed tokens
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = tid + u * FT;
    cid[u] = -1; ck[u] = 0.0f; cb[u] = 0;
    if (j < ncand) {
      cid[u] = cand[j];
      ck[u] = words[cid[u] + mis];
      cb[u] = bucket_of(ck[u], rscale2, b_off2);
      atomicAdd(&hist[cb[u]], 1u);
    }
  }
  __syncthreads();
  {   // the shared-memory row is not read again: the next row's bulk copy starts now, behind the rest of this row
    if (tid == 0 && nrow >= 0) f_issue_row(P, nrow, sc->bar, const_cast<float*>(words), &sc->issued_row);
  }
  {
    uint32_t excl, total2;
    scan_hist(&excl, &total2);
    uint32_t e2 = excl;
#pragma unroll
    for (int b = 0; b < F_BPT; ++b) {
      hist[tid * F_BPT + b] = hloc[b] | ((e2 < 0xffffu ? e2 : 0xffffu) << 16);
      e2 += hloc[b];
    }
    locate_pos(excl, (uint32_t)(K - 1));
  }
  const int tb = sc->sel_bin;
  const int prefix = (int)sc->sel_prefix;                  // tokens in the buckets before tb: all kept
  // (every thread read its list entries before the scan's barriers: the list words may be overwritten now)
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (cid[u] >= 0) {
      if (cb[u] < tb) {
        const uint32_t old = atomicSub(&hist[cb[u]], 1u);    // low 16 bits: slots still free in the bucket
        const uint32_t cnt_left = old & 0xffffu, ex = old >> 16;
        CandEntry e; e.ebits = ex; e.id = cid[u]; e.w = 0u; e.key = ck[u] + 0.0f;
        top[ex + cnt_left - 1u] = e;
      } else if (cb[u] == tb) {
        const int s2 = atomicAdd(&sc->u_n, 1);
        if (s2 < F_RB_CAP) { bnd[s2].id = cid[u]; bnd[s2].kept = __float_as_int(ck[u] + 0.0f); }
      }
    }
  }
  __syncthreads();
  lap(13);
  const int nbnd = sc->u_n;
  if (nbnd > F_RB_CAP) { if (tid == 0) hand_over(P, slow_ws, row, F_WHY_BUCKET); return 1; }
  auto before = [&](float ka, int ia, float kb, int ib) -> bool { return ka > kb || (ka == kb && ia < ib); };   // coder order
  {   // boundary bucket: its first K - prefix tokens in coder order complete the kept set, already in order
    const int need = K - prefix;
    if (tid < nbnd) {
      const float mk = __int_as_float(bnd[tid].kept);
      const int mi = bnd[tid].id;
      int r = 0;
      for (int o = 0; o < nbnd; ++o) r += before(__int_as_float(bnd[o].kept), bnd[o].id, mk, mi) ? 1 : 0;
      if (r < need) { CandEntry e; e.ebits = 0xffffffffu; e.id = mi; e.w = (uint32_t)(prefix + r); e.key = mk; top[prefix + r] = e; }
    }
  }
  __syncthreads();
  // ---- order inside each gathered bucket, exact e of the K kept tokens at their sorted positions
  double* es = reinterpret_cast<double*>(hist);            // [F_K_CAP]
  int* sid = reinterpret_cast<int*>(hist + 2 * F_K_CAP);   // [F_K_CAP]
  int my_r = -1, my_id = 0;
  float my_key = 0.0f;
  if (tid < K) {
    const CandEntry me = top[tid];
    my_id = me.id; my_key = me.key;
    if (me.ebits == 0xffffffffu) my_r = (int)me.w;         // boundary token: position known
    else {
      // its bucket occupies top[ex .. ex + n): n = distance to the next entry with another prefix
      const int ex = (int)me.ebits;
      int r = ex;
      for (int o = ex; o < prefix && top[o].ebits == (uint32_t)ex; ++o) r += before(top[o].key, top[o].id, me.key, me.id) ? 1 : 0;
      my_r = r;
    }
  }
  __syncthreads();                                         // histogram words are free now
  auto a_of = [&](float key) -> double {
    double x = (double)fmaxf(key, clamp_key);
    if (!UNIT_TEMP) x = __ddiv_rn(x, P.temp);
    return x - dm;
  };
  if (tid < K) { es[my_r] = ns_exp64_core(a_of(my_key), tab); sid[my_r] = my_id; }
  __syncthreads();
  lap(14);
  const double ev = tid < K ? es[tid] : 0.0;               // thread r holds the token of rank r
  double S = ev, zero = 0.0;
  u64 none = 0;
  f_sum_ddu(S, zero, none, sc->red);                       // sum of the kept e, fixed order (:146)
  const u64 lo = m_lo, R = m_hi - m_lo;
  const double C = __ddiv_rn((double)R, S);
  const u64 q = tid < K ? (u64)__double2ll_rn(ev * C) : 0ull;   // :146-149
  u64 cum = q;                                             // inclusive prefix sums over the ranks (:150)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, cum, o); if (lane >= o) cum += t; }
  __syncthreads();                                         // red of the sum is consumed
  if (lane == 31) sc->red[warp] = cum;
  __syncthreads();
  u64 Q = 0;
  {
    u64 wo = 0;
#pragma unroll
    for (int w = 0; w < FW; ++w) { const u64 x = sc->red[w]; if (w < warp) wo += x; Q += x; }
    cum += wo;
  }
  u64* cums = reinterpret_cast<u64*>(top);                 // [F_K_CAP]; the gathered list is no longer needed
  if (tid < K) cums[tid] = cum;
  if (tid == 0) { sc->res_idx = K; sc->res_found = 0; }
  __syncthreads();
  // ---- overfill (:153-158): drop the ranks from the first prefix sum above the range on
  int kk = K;
  u64 slack;
  if (Q > R) {
    if (tid < K && cum > R && (tid == 0 || cums[tid - 1] <= R)) sc->res_idx = tid;
    __syncthreads();
    kk = sc->res_idx;
    slack = R - (kk > 0 ? cums[kk - 1] : 0ull);
    __syncthreads();
    if (tid == 0) sc->res_idx = K;
    __syncthreads();
  } else {
    slack = R - Q;
  }
  // bin of rank r: [cums[r-1] + slack, cums[r] + slack), rank 0 starts at 0 and absorbs the slack (:158)
  const u64 my_lo = (tid > 0 && tid < K) ? cums[tid - 1] + slack : 0ull;
  const u64 my_hi = cum + slack;
  if (MODE == MODE_ENC) {
    const u64 m_rel = m_window - lo;                       // next `precision` message bits (:168-171)
    if (tid < kk && my_lo <= m_rel && m_rel < my_hi) sc->res_idx = tid;   // :172 (empty bins never match)
    __syncthreads();
    const int r = sc->res_idx;
    if (tid == (r < kk ? r : 0)) {
      if (r >= kk && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW);     // cannot happen: the bins tile the range
      finish_encode(P, row, m_slot, sid[tid], lo + my_lo, lo + my_hi, (u64)K, Q, m_cursor, m_mlen);   // :175-176
    }
  } else {
    int tok = m_tok;
    if (tok < 0 || tok >= P.V) tok = top_id;
    if (tid < kk && sid[tid] == tok) { sc->res_idx = tid; sc->res_found = 1; }
    __syncthreads();
    const bool in_range = sc->res_found != 0;
    const int r = in_range ? sc->res_idx : 0;              // :342 / :347-348: unknown tokens are coded as rank 0
    if (tid == r) finish_decode(P, row, m_slot, in_range, lo + my_lo, lo + my_hi, (u64)K, Q);
  }
  return 1;
}

template <bool UNIT_TEMP, int MODE, bool RANK>
__device__ __forceinline__ void fast_row(const ns_ac_params& P, int32_t* slow_ws, const int row, const int nrow, const RowMeta meta,
                                         const FastSmem sm, uint32_t& parity, PhaseClock& pc) {
  double* tab = sm.tab; uint32_t* hist = sm.hist; BandEntry* band = sm.band; int* ulist = sm.ulist;
  CandEntry* clist = sm.clist; FScal* sc = sm.sc; float* words = sm.words;
  float4* w4 = reinterpret_cast<float4*>(words);
  // packed-word storage: shared memory, or (stream variant) L2-resident global scratch
  auto word_ld4 = [&](int c) -> float4 {
    return w4[c];
  };
  auto word_st4 = [&](int c, const float4 v) {
    w4[c] = v;
  };
  auto word_ld1 = [&](int i) -> float {
    return words[i];
  };
  const int tid = threadIdx.x;
  const int V = P.V;
  const double temp = P.temp;
  const float c2 = (float)(1.4426950408889634 / temp);     // log2(e)/temp for the fp32 estimate
  const double magic = 6755399441055744.0;                 // 1.5 * 2^52
  // Called by every thread right after a CTA barrier that follows the row's last access to the shared-memory
  // row: thread 0 starts the next row's bulk copy, whose latency then overlaps the rest of this row.
  auto next_row_copy = [&]() {
    if (tid == 0 && nrow >= 0) f_issue_row(P, nrow, sc->bar, words, &sc->issued_row);
  };
  int phase = meta.phase;
  // a row that is skipped still has to consume its bulk copy if the previous row already started it
  auto drain = [&]() {
    if (sc->issued_row == row) {
      const float* g0 = P.logits + (size_t)row * (size_t)P.ld;
      const int NI0 = ((((int)(((uintptr_t)g0 & 15u) >> 2)) + V + 3) >> 2) - 2;
      for (int k = 0; k < F_PIECES; ++k)
        if (NI0 - k * F_SUB * 3 * FT > 0) { f_mbar_wait(&sc->bar[k], (parity >> k) & 1u); parity ^= (1u << k); }
    }
  };
  if (phase == NS_PHASE_DONE) { drain(); return; }
  if (MODE != MODE_ENC) phase = NS_PHASE_CODING;
  const int slot = meta.slot;
  {
    if (MODE == MODE_ENC && P.ntok && slot >= P.token_cap) {
      if (tid == 0) {
        if (P.phase) P.phase[row] = NS_PHASE_DONE;
        if (P.status) atomicOr(&P.status[row], NS_ST_TOKEN_OVERFLOW);
      }
      drain();
      return;
    }
    if (MODE == MODE_DEC && P.ntok_total && slot >= meta.mlen) {
      if (tid == 0 && P.phase) P.phase[row] = NS_PHASE_DONE;
      drain();
      return;
    }

    // ------------------------------------------------------------------ L: bulk copy + estimate
    const float* g = P.logits + (size_t)row * (size_t)P.ld;
    const int mis = (int)(((uintptr_t)g & 15u) >> 2);
    const int W4 = (mis + V + 3) >> 2;                     // float4 chunks of the padded row
    const int NI = W4 - 2;                                 // interior chunks: wholly inside the row
    const int PC = F_SUB * 3 * FT;                         // chunks per piece: every thread does F_SUB x 3 of each piece
    (void)PC;
    if (tid == 0) {
      // (unless the previous row already started this copy once it was done with the buffer)
      if (sc->issued_row != row) f_issue_row(P, row, sc->bar, words, &sc->issued_row);
      sc->band_n = 0; sc->u_n = 0; sc->c_n = 0; sc->bail = 0; sc->band_cut_int = 0; sc->remax = 0; sc->band_kept_n = 0;
    }
    // the two edge chunks may straddle the row ends: plain loads, -inf padding
    if (tid < 8) {
      const int c = tid < 4 ? 0 : W4 - 1;
      const int b = 4 * c - mis + (tid & 3);
      words[4 * c + (tid & 3)] = (b >= 0 && b < V) ? g[b] : -INFINITY;
    }
    auto raw4_l = [&](int c) -> float4 { return w4[c]; };
    auto raw4_p = [&](int c) -> float4 { return w4[c]; };    // masks are written into the shared row after L
    for (int i = tid; i < F_NB / 4; i += FT) reinterpret_cast<uint4*>(hist)[i] = make_uint4(0, 0, 0, 0);
    {   // prefetch this CTA's next row into L2 while this one is processed
      // one bulk prefetch per warp leader: the copy engine walks the lines.  Per-lane prefetch instructions
      // (32 lines each) occupy the load/store pipe for ~1.5k cycles and hold back the estimate's shared loads.
      if (nrow >= 0 && (tid & 31) == 0) {
        const char* np = reinterpret_cast<const char*>(P.logits + (size_t)nrow * (size_t)P.ld);
        const char* a0 = reinterpret_cast<const char*>(((uintptr_t)np + 15u) & ~(uintptr_t)15u);
        const int nbytes = (int)(np + (size_t)V * 4 - a0) & ~15;
        const int per = ((nbytes / FW) + 15) & ~15;
        const int o = (tid >> 5) * per;
        int n = nbytes - o;
        if (n > per) n = per;
        if (n > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0 + o), "r"(n) : "memory");
      }
    }
    pc.mark(0);                                            // row prologue: copy issue, edges, prefetch
    // fp32 online softmax over the pieces as they land: (tm, ts) per thread, lowest id of the max
    float tm = -3.0e38f, ts = 0.f, ntc = 3.0e38f * c2;     // ntc = -tm * c2
    float kmin = 3.0e38f;                                  // lowest logit of the row (bucket range)
    int ti = 0;
    auto online4 = [&](const float4 v, const int b) {
      const float cm = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      kmin = fminf(kmin, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
      if (cm > tm) {
        ts *= f_ex2((tm - cm) * c2);
        tm = cm;
        ntc = -cm * c2;
        ti = (v.x == cm) ? b : (v.y == cm) ? b + 1 : (v.z == cm) ? b + 2 : b + 3;
      }
      ts += (f_ex2(fmaf(v.x, c2, ntc)) + f_ex2(fmaf(v.y, c2, ntc))) + (f_ex2(fmaf(v.z, c2, ntc)) + f_ex2(fmaf(v.w, c2, ntc)));
    };
    auto online12 = [&](const float4 va, const float4 vb, const float4 vc, const int ca, const int cb, const int cc) {
      const float ma = fmaxf(fmaxf(va.x, va.y), fmaxf(va.z, va.w));
      const float mb = fmaxf(fmaxf(vb.x, vb.y), fmaxf(vb.z, vb.w));
      const float mc = fmaxf(fmaxf(vc.x, vc.y), fmaxf(vc.z, vc.w));
      const float cm = fmaxf(ma, fmaxf(mb, mc));
      kmin = fminf(kmin, fminf(fminf(fminf(va.x, va.y), fminf(va.z, va.w)),
                               fminf(fminf(fminf(vb.x, vb.y), fminf(vb.z, vb.w)), fminf(fminf(vc.x, vc.y), fminf(vc.z, vc.w)))));
      if (cm > tm) {                                         // rare after the first pieces
        ts *= f_ex2((tm - cm) * c2);
        tm = cm;
        ntc = -cm * c2;
        const float4 vv = (ma == cm) ? va : (mb == cm) ? vb : vc;
        const int bb = 4 * ((ma == cm) ? ca : (mb == cm) ? cb : cc) - mis;
        ti = (vv.x == cm) ? bb : (vv.y == cm) ? bb + 1 : (vv.z == cm) ? bb + 2 : bb + 3;
      }
      const float sa = (f_ex2(fmaf(va.x, c2, ntc)) + f_ex2(fmaf(va.y, c2, ntc))) + (f_ex2(fmaf(va.z, c2, ntc)) + f_ex2(fmaf(va.w, c2, ntc)));
      const float sb = (f_ex2(fmaf(vb.x, c2, ntc)) + f_ex2(fmaf(vb.y, c2, ntc))) + (f_ex2(fmaf(vb.z, c2, ntc)) + f_ex2(fmaf(vb.w, c2, ntc)));
      const float sc3 = (f_ex2(fmaf(vc.x, c2, ntc)) + f_ex2(fmaf(vc.y, c2, ntc))) + (f_ex2(fmaf(vc.z, c2, ntc)) + f_ex2(fmaf(vc.w, c2, ntc)));
      ts += (sa + sb) + sc3;
    };
    for (int k = 0; k < F_PIECES; ++k) {
      const int c0 = 1 + k * PC;
      int c1 = c0 + PC;
      if (c1 > 1 + NI) c1 = 1 + NI;
      if (c0 < c1) { f_mbar_wait(&sc->bar[k], (parity >> k) & 1u); parity ^= (1u << k); }
      // blocks of 3 float4 per thread: load all three, then run the three dependent chains interleaved
#pragma unroll 1
      for (int sb = 0; sb < F_SUB; ++sb) {
        const int ca = c0 + sb * 3 * FT + tid, cb = ca + FT, cc = cb + FT;
        if (cc < c1) {
          const float4 va = w4[ca], vb = w4[cb], vc = w4[cc];
          online12(va, vb, vc, ca, cb, cc);
        } else {
          if (ca < c1) online4(w4[ca], 4 * ca - mis);
          if (cb < c1) online4(w4[cb], 4 * cb - mis);
        }
      }
    }
    pc.mark(1);                                            // L: waits + estimate over the pieces
    __syncthreads();                                       // edge chunks written by threads 0..7
    {
      const float keep = kmin;                             // the edge chunks carry -inf padding: not part of the range
      if (tid == 0) online4(raw4_l(0), -mis);
      if (tid == 32) online4(raw4_l(W4 - 1), 4 * (W4 - 1) - mis);
      kmin = keep;
      if (tid == 0 || tid == 32) {
        const float4 v = tid == 0 ? raw4_l(0) : raw4_l(W4 - 1);
        if (v.x > -INFINITY) kmin = fminf(kmin, v.x);
        if (v.y > -INFINITY) kmin = fminf(kmin, v.y);
        if (v.z > -INFINITY) kmin = fminf(kmin, v.z);
        if (v.w > -INFINITY) kmin = fminf(kmin, v.w);
      }
    }
    float M, ssum, key_min;
    int top_id;
    if (RANK) {
      // Row reductions with one barrier (the instantiation that carries the rank form, where no fp64 pass follows;
      // in the threshold-only kernel this form makes the exp pass's code slower by more than it saves): warp partials
      // (max key with its lowest id, the sum rescaled to the warp's max, the lowest key) meet in shared memory, then
      // every warp combines the FW partials the same way -- identical bits in every thread, no broadcast.
      float xmask[2] = {-INFINITY, -INFINITY};             // raw logits of the forbidden tokens (estimate's correction)
      const int lane = tid & 31, warp = tid >> 5;
      const uint32_t ok = ns_f32_orderable(tm + 0.0f);
      const uint32_t wk = __reduce_max_sync(0xffffffffu, ok);
      const int wi = __reduce_min_sync(0xffffffffu, ok == wk ? ti : 0x7fffffff);
      const float wm = key_of_pack((u64)wk << 32);
      float wts = ts * f_ex2((tm - wm) * c2);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wts += __shfl_xor_sync(0xffffffffu, wts, o);
      const uint32_t wmin = __reduce_min_sync(0xffffffffu, ns_f32_orderable(kmin));
      uint4* red4 = reinterpret_cast<uint4*>(sc->red);
      if (lane == 0) red4[warp] = make_uint4(wk, (uint32_t)wi, __float_as_uint(wts), wmin);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int id = P.mask_id[k];
        if (id >= 0 && id < V) xmask[k] = words[id + mis];
      }
      __syncthreads();
      const uint4 pr = red4[lane < FW ? lane : 0];
      const uint32_t mk = __reduce_max_sync(0xffffffffu, pr.x);
      top_id = __reduce_min_sync(0xffffffffu, pr.x == mk ? (int)pr.y : 0x7fffffff);
      M = key_of_pack((u64)mk << 32);
      key_min = key_of_pack((u64)__reduce_min_sync(0xffffffffu, pr.w) << 32);
      float part = lane < FW ? __uint_as_float(pr.z) * f_ex2((key_of_pack((u64)pr.x << 32) - M) * c2) : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      ssum = part;
      // forbidden tokens (code_base/arithmetic.py:124-125): probability exactly 0
      bool remax = false;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int id = P.mask_id[k];
        if (id >= 0 && id < V) {
          if (xmask[k] > -INFINITY) ssum -= f_ex2((xmask[k] - M) * c2);
          if (id == top_id) remax = true;
          if (tid == k) words[id + mis] = -INFINITY;
        }
      }
      __syncthreads();                                     // the -inf is visible to whoever sweeps the row next
      if (remax) {                                         // rare: the row maximum itself was forbidden
        u64 pm = 0;
        for (int c = tid; c < W4; c += FT) {
          const float4 v = raw4_p(c);
          const int b = 4 * c - mis;
          u64 p;
          p = pack_of(v.x + 0.0f, b); pm = p > pm ? p : pm;
          p = pack_of(v.y + 0.0f, b + 1); pm = p > pm ? p : pm;
          p = pack_of(v.z + 0.0f, b + 2); pm = p > pm ? p : pm;
          p = pack_of(v.w + 0.0f, b + 3); pm = p > pm ? p : pm;
        }
        pm = f_reduce_u(pm, OpMaxU(), sc->red);
        M = key_of_pack(pm);
        top_id = id_of_pack(pm);
        float s = 0.f;
        for (int c = tid; c < W4; c += FT) {
          const float4 v = raw4_p(c);
          s += f_ex2((v.x - M) * c2) + f_ex2((v.y - M) * c2) + f_ex2((v.z - M) * c2) + f_ex2((v.w - M) * c2);
        }
        ssum = f_sum_f(s, sc->red);
      }
    } else {
      u64 pmax = pack_of(tm + 0.0f, ti), pmin_k = pack_of(kmin, 0);
      f_reduce_maxmin(pmax, pmin_k, sc->red);
      M = key_of_pack(pmax);
      top_id = id_of_pack(pmax);
      ssum = f_sum_f(ts * f_ex2((tm - M) * c2), sc->red);
      // lowest interior logit (-inf if the caller masked tokens with -inf: then the cutoff bounds the range)
      key_min = key_of_pack(pmin_k);
      // forbidden tokens (code_base/arithmetic.py:124-125): probability exactly 0
      if (tid == 0) {
        int remax = 0;
        for (int k = 0; k < 2; ++k) {
          const int id = P.mask_id[k];
          if (id >= 0 && id < V) {
            const float x = words[id + mis];
            if (x > -INFINITY) { ssum -= f_ex2((x - M) * c2); words[id + mis] = -INFINITY; }
            if (id == top_id) remax = 1;
          }
        }
        sc->remax = remax;
        sc->sum32 = ssum;
      }
      __syncthreads();
      if (sc->remax) {                                       // rare: the row maximum itself was forbidden
        u64 pm = 0;
        for (int c = tid; c < W4; c += FT) {
          const float4 v = raw4_p(c);
          const int b = 4 * c - mis;
          u64 p;
          p = pack_of(v.x + 0.0f, b); pm = p > pm ? p : pm;
          p = pack_of(v.y + 0.0f, b + 1); pm = p > pm ? p : pm;
          p = pack_of(v.z + 0.0f, b + 2); pm = p > pm ? p : pm;
          p = pack_of(v.w + 0.0f, b + 3); pm = p > pm ? p : pm;
        }
        pm = f_reduce_u(pm, OpMaxU(), sc->red);
        M = key_of_pack(pm);
        top_id = id_of_pack(pm);
        float s = 0.f;
        for (int c = tid; c < W4; c += FT) {
          const float4 v = raw4_p(c);
          s += f_ex2((v.x - M) * c2) + f_ex2((v.y - M) * c2) + f_ex2((v.z - M) * c2) + f_ex2((v.w - M) * c2);
        }
        s = f_sum_f(s, sc->red);
        if (tid == 0) sc->sum32 = s;
        __syncthreads();
      }
      ssum = sc->sum32;

    }

    if (MODE == MODE_ENC && phase == NS_PHASE_TAIL) {
      if (tid == 0) finish_tail(P, row, slot, top_id);
      return;
    }

    // ------------------------------------------------------------------ row constants
    const u64 lo = meta.lo, hi = meta.hi;
    const u64 R = hi - lo;                                   // arithmetic.py:140
    const double thr = __ddiv_rn(1.0, (double)R);            // :141
    const double Md = (double)M;
    const double dm = UNIT_TEMP ? Md : __ddiv_rn(Md, temp);
    // provisional cutoff from the fp32 estimate: p >= 1/R  <=>  key >= M + temp * ln(sum / R).  fp32 log2 is plenty:
    // the band around the cutoff (F_BAND_EPS) absorbs the error and the split is verified exactly after the exp pass
    float kappa_hi, kappa_lo, clamp_key;
    if (RANK) {                                              // every thread, same bits: no broadcast, no barrier
      const double theta_est = thr * (double)ssum;
      const float tf = (float)temp;
      const float key_th = fmaf(tf * 0.6931471805599453f, __log2f((float)theta_est), M);
      kappa_hi = key_th + tf * F_BAND_EPS; kappa_lo = key_th - tf * F_BAND_EPS;
      clamp_key = (float)(Md - 700.0 * temp);
      if (tid == 0) sc->band_E = ((__double2hiint(theta_est) >> 20) & 0x7ff) - 1024;   // ilogb(theta_est) - 1; read after barriers
      if (!(ssum > 0.0f) || !(R >= 2) || !(kappa_lo > clamp_key)) {
        if (tid == 0) hand_over(P, slow_ws, row, F_WHY_EST);
        return;
      }
    } else {
      if (tid == 0) {
        const double theta_est = thr * (double)ssum;
        int bail = !(ssum > 0.0f) || !(R >= 2);
        const double a_th = (double)(0.6931471805599453f * __log2f((float)theta_est));
        const double key_th = Md + temp * a_th;
        sc->kappa_hi = (float)(key_th + temp * (double)F_BAND_EPS);
        sc->kappa_lo = (float)(key_th - temp * (double)F_BAND_EPS);
        sc->clamp_key = (float)(Md - 700.0 * temp);
        sc->band_E = ilogb(theta_est) - 1;
        if (!(sc->kappa_lo > sc->clamp_key)) bail = 1;
        sc->bail = bail;
      }
      __syncthreads();
      if (sc->bail) { if (tid == 0) hand_over(P, slow_ws, row, F_WHY_EST); return; }
      kappa_hi = sc->kappa_hi; kappa_lo = sc->kappa_lo; clamp_key = sc->clamp_key;
    }

    auto a_of = [&](float key) -> double {                  // (double(x)/temp) - (double(max)/temp), :128-130
      double x = (double)fmaxf(key, clamp_key);
      if (!UNIT_TEMP) x = __ddiv_rn(x, temp);
      return x - dm;
    };
    auto band_push = [&](int id, double e) {
      const int s = atomicAdd(&sc->band_n, 1);
      if (s < F_BAND_CAP) { band[s].id = id; band[s].kept = 0; band[s].e = e; }
    };

    if (RANK && P.topk < V && P.topk >= 2 && P.topk <= F_K_CAP) {
      // top-k binds if more than topk tokens are above the cutoff even should the estimate be 2% off
      const float kappa_r = kappa_hi + 0.02f * (float)temp;
      if (fast_rank_row<UNIT_TEMP, MODE>(P, slow_ws, row, nrow, meta.lo, meta.hi, meta.window, meta.slot, meta.cursor, meta.mlen,
                                        meta.tok, hist, sc, clist, band, words, tab, mis, W4, M, top_id, kappa_r, clamp_key, dm)) {
        pc.mark(7);
        return;
      }
    }
    pc.mark(2);                                            // reductions, masks, row constants
    // ------------------------------------------------------------------ P1: the fp64 exp pass
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    double accl0 = 0.0, accl1 = 0.0, accl2 = 0.0, accl3 = 0.0;
    int cnt_hi = 0;
    const bool need_count = P.topk < V;                      // otherwise only "at least 2 kept" matters
    for (int c = tid; c < W4; c += FT) {
      const float4 v = raw4_p(c);
      const int b = 4 * c - mis;
      const double e0 = ns_exp64_core(a_of(v.x), tab);
      const double e1 = ns_exp64_core(a_of(v.y), tab);
      const double e2 = ns_exp64_core(a_of(v.z), tab);
      const double e3 = ns_exp64_core(a_of(v.w), tab);
      acc0 += e0; acc1 += e1; acc2 += e2; acc3 += e3;
      const bool h0 = v.x >= kappa_hi, h1 = v.y >= kappa_hi, h2 = v.z >= kappa_hi, h3 = v.w >= kappa_hi;
      const bool l0 = v.x < kappa_lo, l1 = v.y < kappa_lo, l2 = v.z < kappa_lo, l3 = v.w < kappa_lo;
      accl0 = __fma_rn(e0, l0 ? 1.0 : 0.0, accl0);           // exact: e * {0,1} + acc
      accl1 = __fma_rn(e1, l1 ? 1.0 : 0.0, accl1);
      accl2 = __fma_rn(e2, l2 ? 1.0 : 0.0, accl2);
      accl3 = __fma_rn(e3, l3 ? 1.0 : 0.0, accl3);
      if (need_count) cnt_hi += (int)h0 + (int)h1 + (int)h2 + (int)h3;
      float4 o;
      o.x = h0 ? f_pack_e(e0) : 0.0f;
      o.y = h1 ? f_pack_e(e1) : 0.0f;
      o.z = h2 ? f_pack_e(e2) : 0.0f;
      o.w = h3 ? f_pack_e(e3) : 0.0f;
      word_st4(c, o);
      if (!((h0 | l0) & (h1 | l1) & (h2 | l2) & (h3 | l3))) {   // rare: inside the guard band
        if (!(h0 | l0)) band_push(b, e0);
        if (!(h1 | l1)) band_push(b + 1, e1);
        if (!(h2 | l2)) band_push(b + 2, e2);
        if (!(h3 | l3)) band_push(b + 3, e3);
      }
    }
    pc.mark(3);                                            // P1 loop
    double sum_all = (acc0 + acc1) + (acc2 + acc3);          // softmax normaliser, :130
    double sum_lo = (accl0 + accl1) + (accl2 + accl3);
    u64 n_hi = (u64)cnt_hi;
    f_sum_ddu(sum_all, sum_lo, n_hi, sc->red);
    const double inv = __ddiv_rn(1.0, sum_all);
    const int nband = sc->band_n;
    const double band_scale = scalbn(1.0, 52 - sc->band_E);
    // ------------------------------------------------------------------ FIX: exact classification
    // band elements, the verification of the provisional split and the bucket range run on different
    // lanes at the same time; one barrier publishes all of it.
    const float kappa_lo_pred = nextafterf(kappa_lo, -INFINITY);
    if (nband <= F_BAND_CAP && tid < nband) {
      const double e = band[tid].e;
      const bool k = (e * inv) >= thr;                       // p_i >= 1/range, :69
      band[tid].kept = k ? 1 : 0;
      if (k) atomicAdd(&sc->band_kept_n, 1);
      else atomicAdd(&sc->band_cut_int, (u64)__double2ull_rz(e * band_scale));   // exact, order-free
    }
    if (tid == FT - 64) {
      // the provisional split is valid iff exp is monotone and both band edges classify as assumed
      const double e_hi = ns_exp64_core(a_of(kappa_hi), tab);
      const double e_lo = ns_exp64_core(a_of(kappa_lo_pred), tab);
      int bail = 0;
      if (nband > F_BAND_CAP) bail = F_WHY_BAND;
      else if (!((e_hi * inv) >= thr) || ((e_lo * inv) >= thr)) bail = F_WHY_VERIFY;
      sc->bail = bail;
    }
    if (tid == FT - 96) {
      // bucket shift: every kept element (certain or band) has e >= e(max(kappa_lo_pred, lowest logit)) > 0
      const float e_min = f_pack_e(ns_exp64_core(a_of(fmaxf(kappa_lo_pred, key_min)), tab));
      const uint32_t span = F_TOP - __float_as_uint(e_min);
      int sh = 0;
      while ((span >> sh) > (uint32_t)(F_NB - 1)) ++sh;
      sc->sh = sh;
    }
    __syncthreads();
    const u64 cand = n_hi + (u64)sc->band_kept_n;            // only counted when topk < V
    const double sum_bc = (double)sc->band_cut_int * scalbn(1.0, sc->band_E - 52);
    const double S = (sum_all - sum_lo) - sum_bc;            // sum of the kept e_i
    // kept set must have 2..topk members, else the reference switches to rank form (:75).  The row
    // maximum has e == 1 exactly and any other kept token has e >= thr * sum_all >= thr, while the
    // rounding error of S is a few ulp of sum_all (< 2^-36 thr-units at precision 31): "another
    // token is kept" <=> S > 1 + thr/2.
    const bool form_ok = need_count ? (cand >= 2 && cand <= (u64)P.topk) : ((inv >= thr) && (S > 1.0 + 0.5 * thr));
    if (sc->bail || !form_ok) {                              // rank form (top-k inside the cutoff set) -> exact kernel
      if (tid == 0) hand_over(P, slow_ws, row, sc->bail ? sc->bail : F_WHY_RANK);
      return;
    }
    const double C = __ddiv_rn((double)R, S);                // :146
    const double C_lo = C * (1.0 - 2.220446049250313e-16);
    const double C_hi = C * (1.0 + 5.960464477539063e-08 + 9.094947017729282e-13);   // e < e_trunc * (1 + 2^-24)
    const int SH = sc->sh;
    auto bin_of_e = [&](float e32) -> uint32_t { return (F_TOP - __float_as_uint(e32)) >> SH; };
    // exact bin width from the original logit (same formula as the exact kernel)
    auto exact_mass = [&](int id) -> uint32_t {
      const float key = g[id] + 0.0f;
      return (uint32_t)__double2ll_rn(ns_exp64_core(a_of(key), tab) * C);
    };
    // bin width decided from the truncated e alone; returns false when it is not decidable
    auto quick_mass = [&](float e32, uint32_t* q) -> bool {
      const double ed = f_unpack_e(e32);
      const uint32_t ql = (uint32_t)ns_double_as_u64(__fma_rn(ed, C_lo, magic));
      const uint32_t qh = (uint32_t)ns_double_as_u64(__fma_rn(ed, C_hi, magic));
      *q = ql;
      return ql == qh;
    };

    pc.mark(4);                                            // FIX: sums, band, verification, constants
    // ------------------------------------------------------------------ decode without a histogram
    // The observed token is known, so the mass ranked before it is a conditional sum over the row: no
    // shared-memory reductions, no bucket scan, no gather.  Tokens whose truncated e equals the token's go to a
    // small list and are ordered exactly (original logit, then id).  Only if the widths overfill the range
    // (arithmetic.py:153-155, about one row in a hundred) the general path below runs instead.
    if (MODE == MODE_DEC) {
      int tok = meta.tok;
      if (tok < 0 || tok >= V) tok = top_id;
      float e32t = word_ld1(tok + mis);
      bool tok_in_band = false;
      if (__float_as_uint(e32t) == 0u) {
        for (int k = 0; k < nband; ++k)
          if (band[k].id == tok && band[k].kept) { e32t = f_pack_e(band[k].e); tok_in_band = true; }
      }
      const uint32_t tbits = __float_as_uint(e32t);          // 0: the token is not in the kept set
      uint32_t qs = 0, bs32 = 0;                             // per-thread sums fit: the whole row's widths are < 2^32
      auto tie_push = [&](int id, uint32_t q) {
        const int s2 = atomicAdd(&sc->c_n, 1);
        if (s2 < F_C_CAP) { clist[s2].ebits = tbits; clist[s2].id = id; clist[s2].w = q; clist[s2].key = 0.0f; }
      };
      auto d2_slow = [&](const float4 v, const int b, const uint32_t* qv, const bool* kv) {
        const float ev[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t bits = __float_as_uint(ev[j]);
          if (!kv[j]) { const int s2 = atomicAdd(&sc->u_n, 1); if (s2 < F_U_CAP) ulist[s2] = b + j; }
          else if (qv[j]) {
            qs += qv[j];
            if (bits > tbits) bs32 += qv[j];
            else if (bits == tbits && b + j != tok) tie_push(b + j, qv[j]);
          }
        }
      };
      for (int c = tid; c < W4; c += 2 * FT) {
        const float4 v = word_ld4(c);
        const float4 w = (c + FT < W4) ? word_ld4(c + FT) : make_float4(0.f, 0.f, 0.f, 0.f);   // packed 0: no-op
        const int b = 4 * c - mis;
        uint32_t q[8];
        bool k[8];
        k[0] = quick_mass(v.x, &q[0]); k[1] = quick_mass(v.y, &q[1]); k[2] = quick_mass(v.z, &q[2]); k[3] = quick_mass(v.w, &q[3]);
        k[4] = quick_mass(w.x, &q[4]); k[5] = quick_mass(w.y, &q[5]); k[6] = quick_mass(w.z, &q[6]); k[7] = quick_mass(w.w, &q[7]);
        const uint32_t bt[8] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w),
                                __float_as_uint(w.x), __float_as_uint(w.y), __float_as_uint(w.z), __float_as_uint(w.w)};
        bool tie = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) tie |= (bt[j] == tbits);
        if ((k[0] & k[1] & k[2] & k[3] & k[4] & k[5] & k[6] & k[7]) && !(tie && tbits != 0u)) {   // e32 == 0 yields q == 0
#pragma unroll
          for (int j = 0; j < 8; ++j) { qs += q[j]; bs32 += bt[j] > tbits ? q[j] : 0u; }
        } else {
          d2_slow(v, b, q, k);
          d2_slow(w, b + 4 * FT, q + 4, k + 4);
        }
      }
      pc.mark(5);
      __syncthreads();
      const int nu = sc->u_n;
      if (nu > F_U_CAP) { if (tid == 0) hand_over(P, slow_ws, row, F_WHY_ULIST); return; }
      for (int u = tid; u < nu; u += FT) {                     // widths the truncated e could not decide
        const int id = ulist[u];
        const uint32_t m = exact_mass(id), bits = __float_as_uint(word_ld1(id + mis));
        qs += m;
        if (bits > tbits) bs32 += m;
        else if (bits == tbits && id != tok && m) tie_push(id, m);
      }
      if (tid < nband && band[tid].kept) {                     // the exact-list band
        const double e = band[tid].e;
        const uint32_t m = (uint32_t)__double2ll_rn(e * C), bits = __float_as_uint(f_pack_e(e));
        qs += m;
        if (bits > tbits) bs32 += m;
        else if (bits == tbits && band[tid].id != tok && m) tie_push(band[tid].id, m);
      }
      __syncthreads();
      const int nt_ties = sc->c_n;
      if (nt_ties > F_C_CAP) { if (tid == 0) hand_over(P, slow_ws, row, F_WHY_BUCKET); return; }
      if (tid < nt_ties) {                                     // same truncated e as the token: original logit, then id
        const float key = g[clist[tid].id] + 0.0f, tkey = g[tok] + 0.0f;
        if (key > tkey || (key == tkey && clist[tid].id < tok)) bs32 += clist[tid].w;
      }
      u64 Qd = (u64)qs, Bd = (u64)bs32;
      {                                                      // exact integer sums: any order
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          Qd += __shfl_xor_sync(0xffffffffu, Qd, o);
          Bd += __shfl_xor_sync(0xffffffffu, Bd, o);
        }
        __syncthreads();
        if ((tid & 31) == 0) { sc->red[tid >> 5] = Qd; sc->red[FW + (tid >> 5)] = Bd; }
        __syncthreads();
        Qd = 0; Bd = 0;
#pragma unroll
        for (int w = 0; w < FW; ++w) { Qd += sc->red[w]; Bd += sc->red[FW + w]; }
      }
      pc.mark(6);
      if (Qd <= R) {
        next_row_copy();                                       // the row buffer is not read again
        const u64 slack = R - Qd;                              // :158
        const u64 top_mass = (u64)__double2ll_rn(C);           // e of the row maximum is exactly 1
        bool in_range = tbits != 0u;
        u64 ws = top_mass, bsum = 0;
        int token = top_id;
        if (in_range) {
          uint32_t qt;
          if (tok_in_band) { for (int k = 0; k < nband; ++k) if (band[k].id == tok) qt = (uint32_t)__double2ll_rn(band[k].e * C); }
          else if (!quick_mass(e32t, &qt)) qt = exact_mass(tok);
          ws = qt; bsum = Bd; token = tok;
        }
        u64 nb, nt;
        if (token == top_id) { nb = lo; nt = lo + ws + slack; }   // :342 / :347-348
        else { nb = lo + bsum + slack; nt = nb + ws; }
        pc.mark(8);
        if (tid == 0) {
          if (P.ntok_total) f_finish_decode(P, row, slot, in_range, nb, nt, cand, Qd, meta.mlen, meta.olen, meta.oword);
          else finish_decode(P, row, slot, in_range, nb, nt, cand, Qd);
        }
        pc.mark(9);
        return;
      }
      // overfill: the general path needs its lists empty again
      __syncthreads();
      if (tid == 0) { sc->u_n = 0; sc->c_n = 0; }
      __syncthreads();
    }
    // ------------------------------------------------------------------ P2: integer bin widths
    // two float4 per iteration: 8 independent conversion + FMA chains in flight per thread
    auto p2_slow = [&](const float4 v, const int b, const uint32_t* qv, const bool* kv) {
      const float ev[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (kv[j]) { if (qv[j]) atomicAdd(&hist[bin_of_e(ev[j])], qv[j]); }
        else { const int s = atomicAdd(&sc->u_n, 1); if (s < F_U_CAP) ulist[s] = b + j; }
      }
    };
    for (int c = tid; c < W4; c += 2 * FT) {
      const float4 v = word_ld4(c);
      const float4 w = (c + FT < W4) ? word_ld4(c + FT) : make_float4(0.f, 0.f, 0.f, 0.f);   // packed 0: no-op
      const int b = 4 * c - mis;
      uint32_t q[8];
      bool k[8];
      k[0] = quick_mass(v.x, &q[0]); k[1] = quick_mass(v.y, &q[1]); k[2] = quick_mass(v.z, &q[2]); k[3] = quick_mass(v.w, &q[3]);
      k[4] = quick_mass(w.x, &q[4]); k[5] = quick_mass(w.y, &q[5]); k[6] = quick_mass(w.z, &q[6]); k[7] = quick_mass(w.w, &q[7]);
      if (k[0] & k[1] & k[2] & k[3] & k[4] & k[5] & k[6] & k[7]) {   // e32 == 0 (not kept) yields q == 0
        f_hist_add(hist, bin_of_e(v.x) & (F_NB - 1), q[0]);  // q == 0: predicated off
        f_hist_add(hist, bin_of_e(v.y) & (F_NB - 1), q[1]);
        f_hist_add(hist, bin_of_e(v.z) & (F_NB - 1), q[2]);
        f_hist_add(hist, bin_of_e(v.w) & (F_NB - 1), q[3]);
        f_hist_add(hist, bin_of_e(w.x) & (F_NB - 1), q[4]);
        f_hist_add(hist, bin_of_e(w.y) & (F_NB - 1), q[5]);
        f_hist_add(hist, bin_of_e(w.z) & (F_NB - 1), q[6]);
        f_hist_add(hist, bin_of_e(w.w) & (F_NB - 1), q[7]);
      } else {
        p2_slow(v, b, q, k);
        p2_slow(w, b + 4 * FT, q + 4, k + 4);
      }
    }
    pc.mark(5);                                            // P2 loop
    __syncthreads();
    const int nu = sc->u_n;
    if (nu > F_U_CAP) { if (tid == 0) hand_over(P, slow_ws, row, F_WHY_ULIST); return; }
    for (int u = tid; u < nu; u += FT) {
      const int id = ulist[u];
      atomicAdd(&hist[bin_of_e(word_ld1(id + mis))], exact_mass(id));
    }
    if (tid < nband && band[tid].kept) {
      const double e = band[tid].e;
      atomicAdd(&hist[bin_of_e(f_pack_e(e))], (uint32_t)__double2ll_rn(e * C));
    }
    __syncthreads();

    // ------------------------------------------------------------------ SEL: bucket scan, kept in registers
    u64 hloc[F_BPT];
    u64 hexcl;                                               // mass in all buckets before this thread's first one
    u64 Q;
    {
      const int lane = tid & 31, warp = tid >> 5;
      u64 tsum = 0;
#pragma unroll
      for (int b = 0; b < F_BPT; ++b) { hloc[b] = hist[tid * F_BPT + b]; tsum += hloc[b]; }
      u64 inc = tsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) sc->red[warp] = inc;
      __syncthreads();
      u64 woff = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < FW; ++w) { const u64 x = sc->red[w]; if (w < warp) woff += x; tot += x; }
      hexcl = woff + inc - tsum;
      Q = tot;
    }
    pc.mark(6);                                            // undecided/band fix-ups, bucket scan
    // first bucket whose inclusive prefix exceeds tau -> sc->sel_bin / sel_prefix
    auto locate = [&](u64 tau) {
      if (tid == 0) { sc->sel_bin = -1; sc->sel_prefix = 0; }
      __syncthreads();
      u64 excl = hexcl;
#pragma unroll
      for (int b = 0; b < F_BPT; ++b) {
        if (hloc[b] != 0 && excl <= tau && tau < excl + hloc[b]) { sc->sel_bin = tid * F_BPT + b; sc->sel_prefix = excl; }
        excl += hloc[b];
      }
      __syncthreads();
    };
    // mass in all buckets before bucket tb -> sc->sel_prefix
    auto prefix_of = [&](int tb) {
      __syncthreads();
      if (tb >= tid * F_BPT && tb < (tid + 1) * F_BPT) {
        u64 excl = hexcl;
        const int off = tb - tid * F_BPT;
#pragma unroll
        for (int b = 0; b < F_BPT; ++b) if (b < off) excl += hloc[b];
        sc->sel_prefix = excl;
      }
      __syncthreads();
    };
    // gather bucket tb: every kept element in it with its exact mass
    auto collect = [&](int tb) -> int {
      if (tid == 0) sc->c_n = 0;
      __syncthreads();
      auto gather4 = [&](const float4 v, const int c) {
        const float ev[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (bin_of_e(ev[j]) == (uint32_t)tb) {             // packed 0 (not kept) maps far outside the histogram
            const int id = 4 * c - mis + j;
            uint32_t q;
            if (!quick_mass(ev[j], &q)) q = exact_mass(id);
            const int s = atomicAdd(&sc->c_n, 1);
            if (s < F_C_CAP) { clist[s].ebits = __float_as_uint(ev[j]); clist[s].id = id; clist[s].w = q; clist[s].key = 0.0f; }
          }
        }
      };
      auto hit4 = [&](const float4 v) -> bool {
        return (bin_of_e(v.x) == (uint32_t)tb) | (bin_of_e(v.y) == (uint32_t)tb) | (bin_of_e(v.z) == (uint32_t)tb) | (bin_of_e(v.w) == (uint32_t)tb);
      };
      // three float4 per iteration: the loads and the twelve bucket tests overlap, matches are rare
      for (int c = tid; c < W4; c += 3 * FT) {
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 va = word_ld4(c);
        const float4 vb = (c + FT < W4) ? word_ld4(c + FT) : zero;
        const float4 vc = (c + 2 * FT < W4) ? word_ld4(c + 2 * FT) : zero;
        const bool ha = hit4(va), hb = hit4(vb), hc = hit4(vc);
        if (ha | hb | hc) {
          if (ha) gather4(va, c);
          if (hb) gather4(vb, c + FT);
          if (hc) gather4(vc, c + 2 * FT);
        }
      }
      if (tid < nband && band[tid].kept) {
        const float e32 = f_pack_e(band[tid].e);
        if (bin_of_e(e32) == (uint32_t)tb) {
          const int s = atomicAdd(&sc->c_n, 1);
          if (s < F_C_CAP) {
            clist[s].ebits = __float_as_uint(e32); clist[s].id = band[tid].id;
            clist[s].w = (uint32_t)__double2ll_rn(band[tid].e * C); clist[s].key = 0.0f;
          }
        }
      }
      __syncthreads();
      int n = sc->c_n;
      if (n > F_C_CAP) return -1;                            // dense bucket: the exact kernel redoes the row
      return n;
    };
    // coder order: larger e first; equal truncated e: larger logit first; equal logits: lower id first
    auto cand_before = [&](const CandEntry& x, const CandEntry& y) -> bool {
      if (x.ebits != y.ebits) return x.ebits > y.ebits;
      if (x.key != y.key) return x.key > y.key;
      return x.id < y.id;
    };
    auto resolve = [&](int n, u64 prefix, bool by_token, u64 tau, int want_id) -> bool {
      // entries sharing a packed e (rare with 24 mantissa bits) need the original logit to be ordered
      for (int c = tid; c < n; c += FT) {
        const uint32_t eb = clist[c].ebits;
        bool d = false;
        for (int o = 0; o < n; ++o) d |= (o != c) && (clist[o].ebits == eb);
        if (d) clist[c].key = g[clist[c].id] + 0.0f;
      }
      if (tid == 0) sc->res_found = 0;
      __syncthreads();                                       // keys and the reset are visible
      for (int c = tid; c < n; c += FT) {
        const CandEntry me = clist[c];
        u64 before = prefix;
        for (int o = 0; o < n; ++o) {
          const CandEntry ot = clist[o];
          if (o != c && cand_before(ot, me)) before += ot.w;
        }
        const bool hit = by_token ? (me.id == want_id) : (me.w != 0 && before <= tau && tau < before + me.w);
        if (hit) { sc->res_idx = me.id; sc->res_before = before; sc->res_w = me.w; sc->res_found = 1; }
      }
      __syncthreads();
      return sc->res_found != 0;
    };
    bool overflow = false;                                   // a gathered bucket did not fit (uniform across the CTA)
    // `last`: nothing reads the shared-memory row after this selection's gather
    auto select_tau = [&](u64 tau, int* idx, u64* before, u64* w, bool last) -> bool {
      locate(tau);
      const int tb = sc->sel_bin;
      const u64 pref = sc->sel_prefix;
      if (tb < 0) { if (last) next_row_copy(); return false; }
      const int n = collect(tb);
      if (last) next_row_copy();
      if (n < 0) { overflow = true; return false; }
      const bool f = resolve(n, pref, false, tau, 0);
      *idx = sc->res_idx; *before = sc->res_before; *w = sc->res_w;
      __syncthreads();
      return f;
    };

    // ------------------------------------------------------------------ overfill (:153-158)
    u64 slack;
    bool truncated = false;
    CandEntry trunc_e = {0u, 0, 0u, 0.0f};
    if (Q > R) {
      int j; u64 bj, wj;
      if (select_tau(R, &j, &bj, &wj, false)) {
        truncated = true;
        trunc_e.ebits = __float_as_uint(word_ld1(j + mis));
        for (int k = 0; k < nband; ++k) if (band[k].id == j) trunc_e.ebits = __float_as_uint(f_pack_e(band[k].e));
        trunc_e.id = j; trunc_e.key = g[j] + 0.0f;
        slack = R - bj;
      } else slack = 0;
    } else {
      slack = R - Q;
    }
    pc.mark(7);                                            // overfill selection
    if (overflow) { if (tid == 0) hand_over(P, slow_ws, row, F_WHY_BUCKET); return; }
    const u64 top_mass = (u64)__double2ll_rn(C);             // e of the row maximum is exactly 1
    u64 nb, nt;
    if (MODE == MODE_ENC) {
      const u64 m_rel = meta.window - lo;                    // next `precision` message bits (:168-171)
      int token;
      if (m_rel < top_mass + slack) {                        // rank 0 absorbs the slack (:158)
        next_row_copy();
        token = top_id; nb = lo; nt = lo + top_mass + slack;
      } else {
        int s; u64 bs, ws;
        if (!select_tau(m_rel - slack, &s, &bs, &ws, true)) {
          if (overflow) { if (tid == 0) hand_over(P, slow_ws, row, F_WHY_BUCKET); return; }
          s = top_id; bs = 0; ws = top_mass;
          if (tid == 0 && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW);
        }
        token = s;                                           // :172
        if (s == top_id) { nb = lo; nt = lo + ws + slack; }
        else { nb = lo + bs + slack; nt = nb + ws; }         // :175-176
      }
      pc.mark(8);                                          // target selection
      if (tid == 0) finish_encode(P, row, slot, token, nb, nt, cand, Q, meta.cursor, meta.mlen);
      pc.mark(9);
    } else {
      int tok = meta.tok;
      if (tok < 0 || tok >= V) tok = top_id;
      // is the observed token in the kept set, and in which bucket?
      float e32t = word_ld1(tok + mis);
      if (__float_as_uint(e32t) == 0u) {
        for (int k = 0; k < nband; ++k)
          if (band[k].id == tok && band[k].kept) e32t = f_pack_e(band[k].e);
      }
      bool in_range = __float_as_uint(e32t) != 0u;
      u64 bs = 0, ws = top_mass;
      int token = top_id;
      if (in_range) {
        const int tb = (int)bin_of_e(e32t);
        prefix_of(tb);
        const u64 pref = sc->sel_prefix;
        const int n = collect(tb);
        if (n < 0) { if (tid == 0) hand_over(P, slow_ws, row, F_WHY_BUCKET); return; }
        if (resolve(n, pref, true, 0, tok)) { bs = sc->res_before; ws = sc->res_w; token = tok; }
        else in_range = false;
        __syncthreads();
        if (in_range && truncated) {
          CandEntry me = {__float_as_uint(e32t), tok, 0u, g[tok] + 0.0f};
          const CandEntry tr = trunc_e;
          if (!cand_before(me, tr)) { in_range = false; token = top_id; bs = 0; ws = top_mass; }
        }
      }
      if (token == top_id) { nb = lo; nt = lo + ws + slack; }   // :342 / :347-348
      else { nb = lo + bs + slack; nt = nb + ws; }
      pc.mark(8);
      if (tid == 0) finish_decode(P, row, slot, in_range, nb, nt, cand, Q);
      pc.mark(9);
    }
  }
}

// RANK: compiled with the rank-form path (host picks it when 2 <= topk <= F_K_CAP and topk < V); the other
// instantiation is the pure threshold-form kernel.  `rows` = nullptr: all P.B rows; else a work list {count, done,
// rows...} (the rows ns_topk.cuh did not carry): the CTAs walk the list and the last one to finish empties it.
template <bool UNIT_TEMP, int MODE, bool RANK = false>
__global__ void __launch_bounds__(FT, NSF_MIN_CTAS) ac_fast_kernel(const __grid_constant__ ns_ac_params P, int32_t* slow_ws, int32_t* rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int n_rows = rows ? rows[0] : P.B;
  if (n_rows > 0) {
    FastSmem sm;
    sm.tab = reinterpret_cast<double*>(smem_raw);
    sm.hist = reinterpret_cast<uint32_t*>(smem_raw + NS_EXP_N * 8);
    sm.band = reinterpret_cast<BandEntry*>(smem_raw + NS_EXP_N * 8 + F_NB * 4);
    sm.ulist = reinterpret_cast<int*>(smem_raw + NS_EXP_N * 8 + F_NB * 4 + F_BAND_CAP * 16);
    sm.clist = reinterpret_cast<CandEntry*>(smem_raw + NS_EXP_N * 8 + F_NB * 4 + F_BAND_CAP * 16 + F_U_CAP * 4);
    sm.sc = reinterpret_cast<FScal*>(smem_raw + NS_EXP_N * 8 + F_NB * 4 + F_BAND_CAP * 16 + F_U_CAP * 4 + 2 * F_C_CAP * 16);
    sm.words = reinterpret_cast<float*>(smem_raw + FIXED_BYTES);   // element id lives at words[id + mis]
    constexpr int HELPER = FT - 32;                          // lane that fetches the next row's scalars
    auto row_at = [&](int i) -> int { return rows ? rows[2 + i] : i; };
    for (int i = tid; i < NS_EXP_N; i += FT) sm.tab[i] = c_exp_tab[i];
    if (tid == 0) {
      for (int k = 0; k < F_PIECES; ++k) f_mbar_init(&sm.sc->bar[k], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      sm.sc->issued_row = -1;
    }
    if (tid == HELPER && (int)blockIdx.x < n_rows) sm.sc->meta[0] = f_load_meta(P, row_at(blockIdx.x), MODE);
    uint32_t parity = 0;                                     // bit k: phase parity of piece k's mbarrier
    PhaseClock pc;
    pc.on = (P.prof != nullptr) && tid == 0;
    pc.last = 0;
    for (int k = 0; k < 16; ++k) pc.acc[k] = 0;
    int it = 0;
    for (int i = blockIdx.x; i < n_rows; i += gridDim.x, ++it) {
      pc.start();
      __syncthreads();                                       // previous row is finished with shared memory
      pc.mark(10);                                           // waiting for the previous row's stragglers
      const RowMeta meta = sm.sc->meta[it & 1];
      RowMeta next;
      const int row = row_at(i);
      const int nrow = i + (int)gridDim.x < n_rows ? row_at(i + (int)gridDim.x) : -1;
      const bool fetch = (tid == HELPER) && (nrow >= 0);
      if (fetch) next = f_load_meta(P, nrow, MODE);          // loads in flight while the row is processed
      fast_row<UNIT_TEMP, MODE, RANK>(P, slow_ws, row, nrow, meta, sm, parity, pc);
      if (fetch) sm.sc->meta[(it + 1) & 1] = next;
      if (pc.on) pc.acc[15] += 1;
    }
    if (pc.on) for (int k = 0; k < 16; ++k) atomicAdd((unsigned long long*)&P.prof[k], (unsigned long long)pc.acc[k]);
  }
  if (rows) {                                                // every CTA has read the count before the last one clears it
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      const int d = atomicAdd(&rows[1], 1);
      if (d == (int)gridDim.x - 1) { rows[0] = 0; rows[1] = 0; __threadfence(); }
    }
  }
}
