// ns_lean.cuh -- arithmetic-coder step, threshold form of the cutoff (code_base/arithmetic.py:127-190), sm_100a.
// Included by ns_coder.cu inside namespace nsl, after the shared definitions (u64, finish_encode, finish_decode, ...).
//
// The throughput kernel of the headline shape (4096 streams x 50257 logits, full distribution).  Same arithmetic and the
// same integers as the exact kernel (ac_step_kernel) -- what is different is the instruction budget: the previous
// kernels (ns_fast.cuh / the two-row variant) executed 76-80 instructions per logit of which 15 were fp64 arithmetic;
// the rest was moves, predicates, branches and bookkeeping the compiler emitted around large unrolled bodies.  Here
// every sweep is one tight loop over the thread's own float4 chunks (chunk c belongs to thread c mod 1024 in every
// sweep, so a thread may patch its own words without a barrier), 32 warps per SM hide the latencies instead of
// unrolling, and everything rare leaves the kernel (hand-over to the exact kernel through slow_ws).
//
// One persistent 1024-thread CTA per SM, the row resident in shared memory (bulk copy issued one row ahead, the row
// after that prefetched into L2).  Per row:
//   L    fp32 estimate: row maximum and sum of 2^((x - ref) c2) against a fixed reference (no online rescaling)
//   P1   ONE fp64 exp per element: exact sum of all e_i (fixed order), exact sum of the provisionally cut ones,
//        elements within 2^-10 of the provisional cutoff to a small exact list; the word is replaced in place by a
//        32-bit truncation of e_i (a per-lane dummy with exponent byte 0 when not kept: width 0, its own bucket)
//   FIX  exact normaliser, provisional cutoff verified, band classified (every warp redundantly: no barrier);
//        C = range / S_kept
//   P2   q_i = rint(e_i C) from the truncated e_i with an interval test (2 DFMA); undecidable ones redone exactly in
//        place.  encode: integer mass histogram (2048 monotone buckets); decode: conditional sums, no histogram
//   SEL  bucket scan -> owner of the target bucket -> gather sweep -> exact order (e32, logit, id) -> mass before
//   UPD  shared-prefix bits, interval rescale, token / bits out (finish_encode / finish_decode)
// Rows the fast path does not carry (finish_sent tail, estimate outside its guard band, rank form, list overflows)
// are queued in slow_ws; the exact kernel then redoes them from scratch with the same integers.

constexpr int LT = 1024;               // threads per CTA
constexpr int LW = LT / 32;            // warps
constexpr int L_NB = 2048;             // histogram buckets
constexpr int L_BAND_CAP = 128;
constexpr int L_C_CAP = 256;           // gathered entries of one bucket
constexpr int L_MIN_VOCAB = 4096;
constexpr float L_BAND_EPS = 0.0009765625f;   // 2^-10 half-width (log units) of the exact-list band
constexpr uint32_t L_TOP = 0xFF000000u;       // packed e of e == 1.0
constexpr uint32_t L_DUMMY_LIMIT = 0x01000000u;   // packed words below this are "not kept" dummies (exponent byte 0)

enum { L_WHY_EST = 1, L_WHY_BAND = 2, L_WHY_VERIFY = 3, L_WHY_RANK = 4, L_WHY_TAIL = 5, L_WHY_BUCKET = 6 };

struct LBand { int id; int pad; double e; };
struct LCand { uint32_t ebits; int id; uint32_t w; float key; };

struct LMeta {
  u64 lo, hi, window;
  int slot, cursor, mlen, tok;
  int phase, olen;
  uint32_t oword, pad;
};

struct LScal {
  LMeta meta[2];
  u64 red[3 * LW];
  u64 bar;
  int band_n, c_n, issued_row, sel_bin, bail, pad_b;
  u64 sel_prefix;
  int res_idx, res_found; u64 res_before; uint32_t res_w, res_ebits; float res_key; int pad2;
  unsigned long long tie_before;
  // per-row constants, written by thread 0 before the exp pass (every thread computes the same bits), read by
  // whoever needs them after the next CTA barrier: nothing has to stay in registers across a sweep
  double k_thr, k_dm; u64 k_R;
  float k_clamp, k_khi, k_klo, k_pad; int k_bandE, k_SH, k_why, k_pad3;
  double k_inv, k_C; u64 k_cand;
};
static_assert(sizeof(LScal) <= 2048, "LScal too large");

// Shared memory: scalars first, then at the next 8 KB boundary of the shared window the histogram (8 KB) and the exp
// table (4 KB) -- a histogram / table address is (index bits) | base, one LOP3 -- then the lists and the row.
// The host requests L_FIXED + (V + 8) * 4 bytes.
constexpr int L_SCAL_BYTES = 2048;
constexpr int L_OFF_TAB = L_NB * 4;                          // offsets from the aligned block: histogram at 0
constexpr int L_OFF_BAND = L_OFF_TAB + NS_EXP_N * 8;
constexpr int L_OFF_CLIST = L_OFF_BAND + L_BAND_CAP * 16;
constexpr int L_OFF_SCAL = L_OFF_CLIST + L_C_CAP * 16;
constexpr int L_OFF_ROW = L_OFF_SCAL + L_SCAL_BYTES;
constexpr int L_FIXED = 8192 + L_OFF_ROW;                    // worst-case alignment gap included
static_assert(L_NB * 4 == 8192, "the histogram fills one 8 KB page");
constexpr int L_MAX_VOCAB = (SMEM_LIMIT - L_FIXED) / 4 - 8;
static_assert(L_MAX_VOCAB >= 50257, "the headline vocabulary must fit");
static_assert(NS_EXP_N * 8 == 4096, "the table fills one 4 KB page");

// constants of the exp core in the constant bank: direct operands of DFMA (no registers, nothing to rematerialise)
__constant__ double c_lk[8] = {NS_512_OVER_LN2, -NS_LN2_512_HI, -NS_LN2_512_LO, 1.0 / 24.0, 1.0 / 6.0, 0.5, 6755399441055744.0, 0.0};

// exp(a) for -708 <= a <= 0: the arithmetic of ns_exp64_core (same bits), with two instructions less around the table.
// The table holds the bit patterns of 2^(j/512) with (j << 11) subtracted from the high word, so adding (n << 11) for
// n = 512 k + j scales by 2^k before the last fma (exact: powers of two commute with rounding, nothing underflows for
// a >= -708).  `tab` = shared-window address of the table, 4 KB aligned.
__device__ __forceinline__ double l_exp64(double a, uint32_t tab) {
  const double magic = 6755399441055744.0;                 // 1.5 * 2^52: low word zero, an immediate operand
  const double t = __fma_rn(a, c_lk[0], magic);
  const uint32_t n = (uint32_t)__double2loint(t);          // low word = rint(a * 512/ln2), two's complement
  const double nd = (double)(int)n;                        // = t - magic exactly; a conversion instead of an fp64-pipe slot
  double r = __fma_rn(nd, c_lk[1], a);
  r = __fma_rn(nd, c_lk[2], r);
  uint32_t tlo, thi;
  asm("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(tlo), "=r"(thi) : "r"(((n << 3) & 0xff8u) | tab));
  double q = __fma_rn(r, c_lk[3], c_lk[4]);
  q = __fma_rn(q, r, 0.5);
  const double r2 = __dmul_rn(r, r);
  const double p = __fma_rn(q, r2, r);
  const double T = __hiloint2double((int)(thi + n * 2048u), (int)tlo);
  return __fma_rn(T, p, T);
}

__device__ __forceinline__ float l_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t l_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void l_mbar_init(u64* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(l_saddr(bar)), "r"(count));
}
__device__ __forceinline__ void l_mbar_expect_tx(u64* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(l_saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l_bulk_g2s(void* dst, const void* src, uint32_t bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(l_saddr(dst)), "l"(src), "r"(bytes), "r"(l_saddr(bar)) : "memory");
}
__device__ __forceinline__ void l_mbar_wait(u64* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" :: "r"(l_saddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t l_pack_e(double e) {
  return __funnelshift_l((uint32_t)__double2loint(e), (uint32_t)__double2hiint(e), 4);
}
__device__ __forceinline__ double l_unpack_e(uint32_t b) {
  return __hiloint2double((int)__funnelshift_r(b, 0x3u, 4), (int)(b << 28));
}
__device__ __forceinline__ void l_hand_over(const ns_ac_params& P, int32_t* slow_ws, int row, int why) {
  const int s = atomicAdd(&slow_ws[0], 1);
  slow_ws[2 + s] = row;
  if (P.status) atomicOr(&P.status[row], NS_ST_EST_RETRY | (why << 8));
}

__device__ __forceinline__ LMeta l_load_meta(const ns_ac_params& P, int row, int mode) {
  LMeta m;
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

// finish_decode with the stream's scalars and its partly filled output word already in registers (stores only)
__device__ __forceinline__ void l_finish_decode(const ns_ac_params& P, int row, int slot, bool in_range, u64 nb, u64 nt,
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

// bin width of one token from its original logit (the exact kernel's formula); rare, so out of line
template <bool UNIT_TEMP>
__device__ __noinline__ uint32_t l_exact_mass(const float* g, int id, float clamp_key, double temp, double dm, double C,
                                              uint32_t tab) {
  double x = (double)fmaxf(g[id] + 0.0f, clamp_key);
  if (!UNIT_TEMP) x = __ddiv_rn(x, temp);
  return (uint32_t)__double2ll_rn(l_exp64(x - dm, tab) * C);
}

// phase timers (thread 0 only); compiled in only for the instantiation the host picks when P.prof is given
template <bool PROF>
struct LClock {
  bool on; long long last; u64 acc[16];
  __device__ __forceinline__ void start() { if (PROF && on) last = clock64(); }
  __device__ __forceinline__ void mark(int k) { if (PROF && on) { const long long t = clock64(); acc[k] += (u64)(t - last); last = t; } }
};

// One opaque register holds the shared-window address of the aligned block; every structure is that plus a constant,
// so no address is ever rebuilt from special registers inside a loop.
struct LSmem {
  uint32_t blk;
  __device__ __forceinline__ unsigned char* at(int off) const { return reinterpret_cast<unsigned char*>(__cvta_shared_to_generic(blk)) + off; }
};

// Start the bulk copy of `row` (interior 16-byte chunks) -- one thread.
__device__ __forceinline__ void l_issue_row(const ns_ac_params& P, int row, LScal* sc, uint32_t* words) {
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  const int NI = ((mis + P.V + 3) >> 2) - 2;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  l_mbar_expect_tx(&sc->bar, (uint32_t)NI * 16u);
  l_bulk_g2s(reinterpret_cast<char*>(words) + 16, reinterpret_cast<const char*>(g - mis) + 16, (uint32_t)NI * 16u, &sc->bar);
  sc->issued_row = row;
}
// The two edge chunks may straddle the row ends: plain loads by 8 lanes, -inf padding outside the row.
__device__ __forceinline__ void l_edges(const ns_ac_params& P, int row, uint32_t* words, int lane8) {
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  const int W4 = (mis + P.V + 3) >> 2;
  const int c = lane8 < 4 ? 0 : W4 - 1;
  const int b = 4 * c - mis + (lane8 & 3);
  words[4 * c + (lane8 & 3)] = __float_as_uint((b >= 0 && b < P.V) ? g[b] : -INFINITY);
}
// L2 prefetch of a whole row, one bulk prefetch per warp leader
__device__ __forceinline__ void l_prefetch_row(const ns_ac_params& P, int row, int warp) {
  const char* np = reinterpret_cast<const char*>(P.logits + (size_t)row * (size_t)P.ld);
  const char* a0 = reinterpret_cast<const char*>(((uintptr_t)np + 15u) & ~(uintptr_t)15u);
  const int nbytes = (int)(np + (size_t)P.V * 4 - a0) & ~15;
  const int per = ((nbytes / LW) + 15) & ~15;
  const int o = warp * per;
  int n = nbytes - o;
  if (n > per) n = per;
  if (n > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0 + o), "r"(n) : "memory");
}

// P1: the fp64 exp pass over the thread's chunks.  Everything the loop touches is an argument.
template <bool UNIT_TEMP>
__device__ __forceinline__ void lean_p1(uint4* w4, int W4, int tid, uint32_t tab, double dm, double temp, float clamp_key,
                                        float kappa_hi, float kappa_lo, uint32_t dummy, LScal* sc, LBand* band, int mis,
                                        double& acc_out, double& accl_out) {
  double acc = 0.0, accl = 0.0;
  // loop constants as opaque register values: rematerialising them inside the loop costs more than holding them
  asm volatile("" : "+f"(clamp_key), "+r"(tab), "+r"(dummy), "+r"(W4));
  auto a_of = [&](float key) -> double {                    // (double(x)/temp) - (double(max)/temp), :128-130
    double x = (double)fmaxf(key, clamp_key);
    if (!UNIT_TEMP) x = __ddiv_rn(x, temp);
    return __dsub_rn(x, dm);
  };
#pragma unroll 1
  for (int c = tid; c < W4; c += LT) {
    const uint4 u = w4[c];
    const float x[4] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)};
    uint32_t o[4];
    bool maybe = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double e = l_exp64(a_of(x[j]), tab);
      acc = __dadd_rn(acc, e);
      // accl = sum of everything not certainly kept (the guard band included): one compare serves the sum, the packing
      // and, with a second one folded onto it, the band test
      const bool hi = x[j] >= kappa_hi;
      if (!hi) accl = __dadd_rn(accl, e);                    // one predicated add
      o[j] = hi ? l_pack_e(e) : dummy;
      maybe |= (!hi) & (x[j] >= kappa_lo);
    }
    w4[c] = make_uint4(o[0], o[1], o[2], o[3]);
    if (maybe) {                                             // rare: inside the guard band
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (!(x[j] >= kappa_hi) && !(x[j] < kappa_lo)) {
          const int s = atomicAdd(&sc->band_n, 1);
          if (s < L_BAND_CAP) { band[s].id = 4 * c - mis + j; band[s].pad = 0; band[s].e = l_exp64(a_of(x[j]), tab); }
        }
    }
  }
  acc_out = acc; accl_out = accl;
}

// bin widths of one chunk from the truncated e: q = rint(e C) is proven when both ends of the interval round alike.
// Returns the mask of lanes (bit j) that are not decidable; their width is set to 0 and settled after the sweep.
__device__ __forceinline__ uint32_t lean_widths(const uint4 u, double C_lo, double C_hi, uint32_t (&q)[4]) {
  const double magic = 6755399441055744.0;                 // 1.5 * 2^52
  const uint32_t bt[4] = {u.x, u.y, u.z, u.w};
  uint32_t bad = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double ed = l_unpack_e(bt[j]);
    const uint32_t ql = (uint32_t)__double2loint(__fma_rn(ed, C_lo, magic));
    const uint32_t qh = (uint32_t)__double2loint(__fma_rn(ed, C_hi, magic));
    q[j] = ql;
    if (ql != qh) bad |= 1u << j;
  }
  return bad;
}

template <bool UNIT_TEMP, int MODE, bool PROF>
__device__ __forceinline__ void lean_row(const ns_ac_params& P, int32_t* slow_ws, const int row, const LMeta* mp,
                                         const LSmem& sm, uint32_t& parity, LClock<PROF>& pc) {
  uint32_t* hist = reinterpret_cast<uint32_t*>(sm.at(0));
  LBand* band = reinterpret_cast<LBand*>(sm.at(L_OFF_BAND));
  LCand* clist = reinterpret_cast<LCand*>(sm.at(L_OFF_CLIST));
  LScal* sc = reinterpret_cast<LScal*>(sm.at(L_OFF_SCAL));
  uint32_t* words = reinterpret_cast<uint32_t*>(sm.at(L_OFF_ROW));
  const uint32_t tab = sm.blk + L_OFF_TAB;
  int tid = threadIdx.x;
  const int V = P.V;
  const double magic = 6755399441055744.0;                 // 1.5 * 2^52
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  int mis = (int)(((uintptr_t)g & 15u) >> 2);
  int W4 = (mis + V + 3) >> 2;
  asm volatile("" : "+r"(tid), "+r"(mis), "+r"(W4));       // held, not recomputed from special registers / parameters
  const int lane = tid & 31, warp = tid >> 5;
  uint4* w4 = reinterpret_cast<uint4*>(words);

  // Called by every thread once nothing reads the row buffer any more (after a CTA barrier): the next row's copy
  // and edge chunks start now, so their latency overlaps the rest of this row.
  auto next_row_copy = [&]() {
    const int nrow = row + (int)gridDim.x;
    if (nrow < P.B) {
      if (tid == 0) l_issue_row(P, nrow, sc, words);
      if (tid >= 32 && tid < 40) l_edges(P, nrow, words, tid - 32);
    }
  };
  auto wait_row = [&]() { l_mbar_wait(&sc->bar, parity); parity ^= 1u; };
  // leave the row: the copy is already in flight (issued by the previous row) and must be consumed
  auto skip_row = [&]() { wait_row(); __syncthreads(); next_row_copy(); };
  // give the row to the exact kernel; every thread calls it at the same point, nobody reads the row buffer afterwards
  auto give_up = [&](int why) {
    __syncthreads();
    if (tid == 0) l_hand_over(P, slow_ws, row, why);
    next_row_copy();
  };
  auto exact_mass = [&](int id, double C) -> uint32_t {
    return l_exact_mass<UNIT_TEMP>(g, id, sc->k_clamp, P.temp, sc->k_dm, C, tab);
  };

  int phase = mp->phase;
  const int slot = mp->slot;
  if (phase == NS_PHASE_DONE) { skip_row(); return; }
  if (MODE != MODE_ENC) phase = NS_PHASE_CODING;
  if (MODE == MODE_ENC && P.ntok && slot >= P.token_cap) {
    if (tid == 0) {
      if (P.phase) P.phase[row] = NS_PHASE_DONE;
      if (P.status) atomicOr(&P.status[row], NS_ST_TOKEN_OVERFLOW);
    }
    skip_row();
    return;
  }
  if (MODE == MODE_DEC && P.ntok_total && slot >= mp->mlen) {
    if (tid == 0 && P.phase) P.phase[row] = NS_PHASE_DONE;
    skip_row();
    return;
  }
  if (MODE == MODE_ENC && phase == NS_PHASE_TAIL) {          // finish_sent tail (:135-137): the exact kernel emits rank 0
    if (tid == 0) l_hand_over(P, slow_ws, row, L_WHY_TAIL);
    skip_row();
    return;
  }

  // ------------------------------------------------------------------ L: the row arrives, fp32 estimate
  hist[tid] = 0u; hist[tid + LT] = 0u;
  if (tid == 0) { sc->band_n = 0; sc->c_n = 0; sc->res_found = 0; sc->tie_before = 0ull; sc->bail = 0; }
  {
    const int nrow = row + (int)gridDim.x;
    if (nrow < P.B && warp == 2) l_prefetch_row(P, nrow, lane);   // one warp, one slice per lane
  }
  pc.mark(0);
  wait_row();
  // forbidden tokens (code_base/arithmetic.py:124-125): probability exactly 0.  The owner of the chunk patches the
  // word before its own first sweep -- no other thread reads it before the next CTA barrier.
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int id = P.mask_id[k];
    if (id >= 0 && id < V && (((id + mis) >> 2) & (LT - 1)) == tid) words[id + mis] = 0xFF800000u;   // -inf
  }
  const float c2 = (float)(1.4426950408889634 / P.temp);   // log2(e)/temp
  float ref;
  {
    // reference of the estimate: the largest of the row's first four logits (any finite value works; the sum is
    // rescaled to the true maximum afterwards, an overflow hands the row over)
    ref = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k != P.mask_id[0] && k != P.mask_id[1]) ref = fmaxf(ref, __uint_as_float(words[mis + k]));
    if (!(ref > -3.0e38f) || !(ref < 3.0e38f)) ref = 0.f;
  }
  float tm = -INFINITY, ts0 = 0.f, ts1 = 0.f;
  {
    const float nrc = -ref * c2;
#pragma unroll 4
    for (int c = tid; c < W4; c += LT) {
      const uint4 u = w4[c];
      const float x0 = __uint_as_float(u.x), x1 = __uint_as_float(u.y), x2 = __uint_as_float(u.z), x3 = __uint_as_float(u.w);
      tm = fmaxf(fmaxf(x0, x1), tm);
      tm = fmaxf(fmaxf(x2, x3), tm);
      ts0 += l_ex2(fmaf(x0, c2, nrc)); ts1 += l_ex2(fmaf(x1, c2, nrc));
      ts0 += l_ex2(fmaf(x2, c2, nrc)); ts1 += l_ex2(fmaf(x3, c2, nrc));
    }
  }
  pc.mark(1);
  float M, ssum;
  {
    const uint32_t ok = ns_f32_orderable(tm + 0.0f);
    const uint32_t wk = __reduce_max_sync(0xffffffffu, ok);
    float wts = ts0 + ts1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wts += __shfl_xor_sync(0xffffffffu, wts, o);
    if (lane == 0) sc->red[warp] = ((u64)wk << 32) | (u64)__float_as_uint(wts);
    __syncthreads();
    const u64 pr = sc->red[lane];
    const uint32_t mk = __reduce_max_sync(0xffffffffu, (uint32_t)(pr >> 32));
    M = key_of_pack((u64)mk << 32);
    float part = __uint_as_float((uint32_t)pr);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    ssum = part * l_ex2((ref - M) * c2);
  }

  // ------------------------------------------------------------------ row constants (every thread, same bits)
  double acc, accl;
  const bool need_count = P.topk < V;                        // otherwise only "at least 2 kept" matters
  const uint32_t dummy = (uint32_t)lane << 18;               // "not kept": exponent byte 0 (width 0), one bucket per lane
  {
    const u64 R = mp->hi - mp->lo;                           // arithmetic.py:140
    const double thr = __drcp_rn((double)R);                 // :141 (correctly rounded, = 1.0 / R)
    const double Md = (double)M;
    const double dm = UNIT_TEMP ? Md : __ddiv_rn(Md, P.temp);
    // provisional cutoff from the fp32 estimate: p >= 1/R  <=>  key >= M + temp * ln(sum / R).  The band around it
    // absorbs the estimate's error and the split is verified exactly after the exp pass.
    const double theta_est = thr * (double)ssum;
    const float tf = (float)P.temp;
    const float key_th = fmaf(tf * 0.6931471805599453f, __log2f((float)theta_est), M);
    const float kappa_hi = key_th + tf * L_BAND_EPS, kappa_lo = key_th - tf * L_BAND_EPS;
    const float clamp_key = (float)(Md - 700.0 * P.temp);
    if (!(ssum > 0.0f) || !(ssum < 3.0e38f) || !(R >= 2) || !(kappa_lo > clamp_key) || !(M > -3.0e38f) || kappa_lo == 0.f) {
      give_up(L_WHY_EST);
      return;
    }
    if (tid == 0) {
      sc->k_thr = thr; sc->k_dm = dm; sc->k_R = R; sc->k_clamp = clamp_key; sc->k_khi = kappa_hi; sc->k_klo = kappa_lo;
      sc->k_bandE = ((__double2hiint(theta_est) >> 20) & 0x7ff) - 1024;   // ilogb(theta_est) - 1
    }
    pc.mark(2);
    // ---------------------------------------------------------------- P1: the fp64 exp pass
    lean_p1<UNIT_TEMP>(w4, W4, tid, tab, dm, P.temp, clamp_key, kappa_hi, kappa_lo, dummy, sc, band, mis, acc, accl);
  }
  int cnt_hi = 0;
  if (need_count) {                                          // 512 < topk < V: count the certainly kept tokens
    for (int c = tid; c < W4; c += LT) {
      const uint4 u = w4[c];
      cnt_hi += (int)(u.x >= L_DUMMY_LIMIT) + (int)(u.y >= L_DUMMY_LIMIT) + (int)(u.z >= L_DUMMY_LIMIT) + (int)(u.w >= L_DUMMY_LIMIT);
    }
  }
  pc.mark(3);
  // exact sums in a fixed order: butterfly inside the warp, warp partials one per lane, second butterfly
  double sum_all, sum_lo;
  u64 n_hi;
  {
    double a = acc, b = accl;
    int cn = cnt_hi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = a + __shfl_xor_sync(0xffffffffu, a, o);
      b = b + __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (need_count) cn = __reduce_add_sync(0xffffffffu, cn);
    if (lane == 0) {
      sc->red[warp] = (u64)__double_as_longlong(a);
      sc->red[LW + warp] = (u64)__double_as_longlong(b);
      sc->red[2 * LW + warp] = (u64)(uint32_t)cn;
    }
    __syncthreads();
    sum_all = 0.0; sum_lo = 0.0; n_hi = 0;
    if (warp < 2) {                                          // only the two warps of FIX need the totals
      a = __longlong_as_double((long long)sc->red[lane]);
      b = __longlong_as_double((long long)sc->red[LW + lane]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a = a + __shfl_xor_sync(0xffffffffu, a, o);
        b = b + __shfl_xor_sync(0xffffffffu, b, o);
      }
      sum_all = a; sum_lo = b;
      n_hi = need_count ? (u64)__reduce_add_sync(0xffffffffu, (int)(uint32_t)sc->red[2 * LW + lane]) : 0ull;
    }
  }
  // ------------------------------------------------------------------ FIX: exact classification
  // One warp derives the row's constants from the exact sums and publishes them; a second one checks the provisional
  // split (its flag is read after the sweep that follows -- nothing irrevocable happens before).  Done by every
  // thread this cost 250 instructions x 32 warps per row.
  const int nband = sc->band_n;
  if (warp == 0) {
    const double inv = __drcp_rn(sum_all);                   // correctly rounded, = 1.0 / sum_all
    const u64 R = sc->k_R;
    const double thr = sc->k_thr;
    u64 band_cut_int = 0;
    int band_kept_n = 0;
    const int band_E = sc->k_bandE;
    int why = 0;
    if (nband > L_BAND_CAP) why = L_WHY_BAND;
    else if (nband > 0) {
      const double band_scale = __hiloint2double((1023 + 52 - band_E) << 20, 0);   // 2^(52 - band_E)
      for (int k = lane; k < nband; k += 32) {
        const double e = band[k].e;
        if ((e * inv) >= thr) { band_kept_n += 1; band_cut_int += (u64)__double2ull_rz(e * band_scale); }   // the KEPT ones here
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        band_cut_int += __shfl_xor_sync(0xffffffffu, band_cut_int, o);
        band_kept_n += __shfl_xor_sync(0xffffffffu, band_kept_n, o);
      }
    }
    const u64 cand = n_hi + (u64)band_kept_n;                // only counted when topk < V
    const double sum_bc = (double)band_cut_int * __hiloint2double((1023 - 52 + band_E) << 20, 0);
    const double S = (sum_all - sum_lo) + sum_bc;            // sum of the kept e_i (sum_lo holds the whole band)
    // kept set must have 2..topk members, else the reference switches to rank form (:75).  The row maximum has
    // e == 1 exactly and any other kept token has e >= thr * sum_all >= thr: "another token is kept" <=> S > 1 + thr/2.
    const bool form_ok = need_count ? (cand >= 2 && cand <= (u64)P.topk) : ((inv >= thr) && (S > 1.0 + 0.5 * thr));
    if (!form_ok && why == 0) why = L_WHY_RANK;
    const double C = __ddiv_rn((double)R, S);                // :146
    // bucket of a packed e: kept elements have e >= 1/R, i.e. at most ilog2(R) + 2 octaves below 1.0
    int SH = 13;
    {
      const uint32_t span = (uint32_t)(66 - __clzll((long long)R)) << 24;
      while ((span >> SH) > (uint32_t)(L_NB - 1)) ++SH;
    }
    if (lane == 0) {
      sc->k_inv = inv; sc->k_C = C; sc->k_cand = cand; sc->k_SH = SH; sc->k_why = why;
    }
  } else if (warp == 1) {
    // the provisional split is valid iff exp is monotone and both band edges classify as assumed
    const double inv = __drcp_rn(sum_all);
    const double thr = sc->k_thr;
    const float kappa_hi = sc->k_khi, kappa_lo = sc->k_klo, clamp_key = sc->k_clamp;
    const double dm = sc->k_dm;
    auto a_of = [&](float key) -> double {
      double x = (double)fmaxf(key, clamp_key);
      if (!UNIT_TEMP) x = __ddiv_rn(x, P.temp);
      return __dsub_rn(x, dm);
    };
    const float kappa_lo_pred = __uint_as_float(__float_as_uint(kappa_lo) + (kappa_lo > 0.f ? 0xffffffffu : 1u));   // next below (kappa_lo != 0)
    const double e_hi = l_exp64(a_of(kappa_hi), tab);
    const double e_lo = l_exp64(a_of(kappa_lo_pred), tab);
    if ((!((e_hi * inv) >= thr) || ((e_lo * inv) >= thr)) && lane == 0) sc->bail = L_WHY_VERIFY;
  }
  __syncthreads();
  if (sc->k_why) { const int why = sc->k_why; give_up(why); return; }
  const u64 R = sc->k_R;
  const double thr = sc->k_thr, inv = sc->k_inv;
  const double C = sc->k_C;
  const double C_lo = C * (1.0 - 2.220446049250313e-16);
  const double C_hi = C * (1.0 + 5.960464477539063e-08 + 9.094947017729282e-13);   // e < e_trunc * (1 + 2^-24)
  const int SH = sc->k_SH;
  const u64 cand = sc->k_cand;
  auto band_kept = [&](int k) -> bool { return (band[k].e * inv) >= thr; };
  pc.mark(4);

  u64 Q = 0;                                                 // total mass of the kept bins
  u64 dec_before = 0, dec_w = 0;                             // decode: mass ranked before the observed token, its width
  bool dec_in_range = false, dec_flag_ok = false;
  uint32_t tbits = 0;
  const int tok = mp->tok;
  bool need_hist = (MODE == MODE_ENC);

  if (MODE == MODE_DEC) {
    // ---------------------------------------------------------------- decode: conditional sums, no histogram
    const bool tok_ok = tok >= 0 && tok < V;
    int tok_band = -1;
    if (tok_ok) {
      tbits = words[tok + mis];
      if (tbits < L_DUMMY_LIMIT) {
        tbits = 0;
        for (int k = 0; k < nband; ++k)
          if (band[k].id == tok && band_kept(k)) { tbits = l_pack_e(band[k].e); tok_band = k; }
      }
    }
    uint32_t qs = 0, bs = 0;
    int pend0 = -1, pend1 = -1;                              // undecidable widths, settled after the sweep
    bool pend_over = false;
    {
      int W4o = W4; uint32_t tb = tbits; double clo = C_lo, chi = C_hi;
      asm volatile("" : "+r"(W4o), "+r"(tb), "+d"(clo), "+d"(chi));
#pragma unroll 1
      for (int c = tid; c < W4o; c += LT) {
        const uint4 u = w4[c];
        const uint32_t bt[4] = {u.x, u.y, u.z, u.w};
        uint32_t q[4];
        const uint32_t bad = lean_widths(u, clo, chi, q);
        bool tie = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) { qs += q[j]; bs += bt[j] > tb ? q[j] : 0u; tie |= bt[j] == tb; }
        if (bad | (uint32_t)tie) {                           // rare
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int id = 4 * c - mis + j;
            if (bad & (1u << j)) {                           // this width was added above: take it out again
              qs -= q[j]; if (bt[j] > tb) bs -= q[j];
              if (pend1 >= 0) pend_over = true;
              pend1 = pend0; pend0 = id;
            } else if (tb != 0u && bt[j] == tb && id != tok && q[j] != 0u) {
              // same truncated e as the observed token: ordered exactly below
              const int s2 = atomicAdd(&sc->c_n, 1);
              if (s2 < L_C_CAP) { clist[s2].ebits = tb; clist[s2].id = id; clist[s2].w = q[j]; clist[s2].key = 0.f; }
            }
          }
        }
      }
    }
    if (pend0 >= 0) {
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int id = t == 0 ? pend0 : pend1;
        if (id < 0) continue;
        const uint32_t m = exact_mass(id, C), bits = words[id + mis];
        qs += m;
        if (bits > tbits) bs += m;
        else if (tbits != 0u && bits == tbits && id != tok && m != 0u) {
          const int s2 = atomicAdd(&sc->c_n, 1);
          if (s2 < L_C_CAP) { clist[s2].ebits = tbits; clist[s2].id = id; clist[s2].w = m; clist[s2].key = 0.f; }
        }
      }
    }
    if (pend_over) sc->bail = L_WHY_BUCKET;
    if (tid < nband && band_kept(tid)) {                     // the exact-list band
      const double e = band[tid].e;
      const uint32_t m = (uint32_t)__double2ll_rn(e * C), bits = l_pack_e(e);
      qs += m;
      if (bits > tbits) bs += m;
      else if (bits == tbits && band[tid].id != tok && m && tbits != 0u) {
        const int s2 = atomicAdd(&sc->c_n, 1);
        if (s2 < L_C_CAP) { clist[s2].ebits = tbits; clist[s2].id = band[tid].id; clist[s2].w = m; clist[s2].key = 0.f; }
      }
    }
    pc.mark(5);
    // the whole row's widths sum to about the range (< 2^32 at precision <= 31): 32-bit partial sums are exact
    u64 Qd = (u64)__reduce_add_sync(0xffffffffu, qs);
    u64 Bd = (u64)__reduce_add_sync(0xffffffffu, bs);
    if (lane == 0) { sc->red[warp] = Qd; sc->red[LW + warp] = Bd; }
    __syncthreads();
    Qd = sc->red[lane]; Bd = sc->red[LW + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Qd += __shfl_xor_sync(0xffffffffu, Qd, o);
      Bd += __shfl_xor_sync(0xffffffffu, Bd, o);
    }
    const int nties = sc->c_n;
    if (nties > L_C_CAP || sc->bail) { const int why = sc->bail ? sc->bail : L_WHY_BUCKET; give_up(why); return; }
    if (nties > 0) {                                         // same truncated e as the token: original logit, then id
      if (tid < nties) {
        const float key = g[clist[tid].id] + 0.0f, tkey = g[tok] + 0.0f;
        if (key > tkey || (key == tkey && clist[tid].id < tok)) atomicAdd(&sc->tie_before, (unsigned long long)clist[tid].w);
      }
      __syncthreads();
      Bd += sc->tie_before;
    }
    Q = Qd;
    dec_in_range = tbits != 0u;
    dec_flag_ok = dec_in_range || !tok_ok;                   // an invalid id is coded as rank 0 (the exact kernel does the same)
    if (dec_in_range) {
      uint32_t qt;
      if (tok_band >= 0) qt = (uint32_t)__double2ll_rn(band[tok_band].e * C);
      else {
        const double ed = l_unpack_e(tbits);
        qt = (uint32_t)__double2loint(__fma_rn(ed, C_lo, magic));
        if (qt != (uint32_t)__double2loint(__fma_rn(ed, C_hi, magic))) qt = exact_mass(tok, C);
      }
      dec_w = qt; dec_before = Bd;
    }
    pc.mark(6);
    if (Q <= R) {
      __syncthreads();                                       // every thread is done with the row buffer and the lists
      next_row_copy();
      const u64 slack = R - Q;                               // :158
      const u64 lo = mp->lo;
      const u64 top_mass = (u64)__double2ll_rn(C);           // e of the row maximum is exactly 1
      u64 nb, nt;
      if (!dec_in_range || dec_before == 0) { nb = lo; nt = lo + (dec_in_range ? dec_w : top_mass) + slack; }   // :342 / :347-348
      else { nb = lo + dec_before + slack; nt = nb + dec_w; }
      pc.mark(8);
      if (tid == 0) {
        if (P.ntok_total) l_finish_decode(P, row, slot, dec_flag_ok, nb, nt, cand, Q, mp->mlen, mp->olen, mp->oword);
        else finish_decode(P, row, slot, dec_flag_ok, nb, nt, cand, Q);
      }
      pc.mark(9);
      return;
    }
    // the widths overfill the range (:153-155, about one row in a hundred): the truncation point needs the histogram
    need_hist = true;
    __syncthreads();
    if (tid == 0) sc->c_n = 0;
  }

  // ------------------------------------------------------------------ P2: integer bin widths -> mass histogram
  if (need_hist) {
    int pend0 = -1, pend1 = -1;                              // undecidable widths, settled after the sweep
    bool pend_over = false;
    {
      int W4o = W4; double clo = C_lo, chi = C_hi;
      uint32_t hb = l_saddr(hist), sh2 = (uint32_t)(SH - 2);
      asm volatile("" : "+r"(W4o), "+d"(clo), "+d"(chi), "+r"(hb), "+r"(sh2));
#pragma unroll 1
      for (int c = tid; c < W4o; c += LT) {
        const uint4 u = w4[c];
        const uint32_t bt[4] = {u.x, u.y, u.z, u.w};
        uint32_t q[4];
        const uint32_t bad = lean_widths(u, clo, chi, q);
        if (bad) {                                           // rare: width 0 for now
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (bad & (1u << j)) {
              q[j] = 0u;
              if (pend1 >= 0) pend_over = true;
              pend1 = pend0; pend0 = 4 * c - mis + j;
            }
        }
        // dummies (width 0) land in 32 per-lane buckets of the wrapped index: no same-address serialisation
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hb | ((bt[j] >> sh2) & ((uint32_t)(L_NB - 1) << 2))), "r"(q[j]) : "memory");
        }
      }
    }
    if (pend0 >= 0) {
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int id = t == 0 ? pend0 : pend1;
        if (id < 0) continue;
        atomicAdd(&hist[(words[id + mis] >> SH) & (L_NB - 1)], exact_mass(id, C));
      }
    }
    if (pend_over) sc->bail = L_WHY_BUCKET;
    if (tid < nband && band_kept(tid)) {
      const double e = band[tid].e;
      atomicAdd(&hist[(l_pack_e(e) >> SH) & (L_NB - 1)], (uint32_t)__double2ll_rn(e * C));
    }
  }
  pc.mark(5);
  __syncthreads();
  if (sc->bail) { const int why = sc->bail; give_up(why); return; }   // the split did not verify (warp 1) / too many undecidable widths

  // ------------------------------------------------------------------ SEL: bucket scan (two buckets per thread)
  // Bucket of a packed e: (bits >> SH) mod 2048 -- no subtraction in the sweep; the kept keys span fewer than 2048
  // consecutive values of bits >> SH below top_v = L_TOP >> SH, so the residues are distinct.  The coder's order is
  // descending keys: position p of that order is bucket (top_v - p) mod 2048; thread t scans positions 2t, 2t + 1.
  const uint32_t top_v = L_TOP >> SH;
  const uint32_t h0 = hist[(top_v - 2u * (uint32_t)tid) & (L_NB - 1)], h1 = hist[(top_v - 2u * (uint32_t)tid - 1u) & (L_NB - 1)];
  uint32_t hexcl;                                            // mass in all buckets before this thread's first one
  {
    const uint32_t tsum = h0 + h1;
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sc->red[warp] = (u64)inc;
    __syncthreads();
    const uint32_t wt = (uint32_t)sc->red[lane];
    uint32_t winc = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    const uint32_t tot = __shfl_sync(0xffffffffu, winc, 31);
    const uint32_t wexcl = __shfl_sync(0xffffffffu, winc - wt, warp);
    hexcl = wexcl + inc - tsum;
    if (MODE == MODE_ENC) Q = (u64)tot;
  }
  pc.mark(6);
  // position tau of the coder's order: bucket owner -> gather sweep -> exact order inside the bucket.
  // Returns false if no kept bin holds tau.  Result in sc->res_* (id, mass before, width, packed e).
  bool overflow = false;
  auto select_tau = [&](u64 tau, bool last) -> bool {
    {
      const u64 e0 = (u64)hexcl, e1 = e0 + h0;
      int b = -1; u64 pre = 0;
      if (h0 != 0u && e0 <= tau && tau < e0 + h0) { b = 2 * tid; pre = e0; }
      if (h1 != 0u && e1 <= tau && tau < e1 + h1) { b = 2 * tid + 1; pre = e1; }
      if (b >= 0) { sc->sel_bin = b; sc->sel_prefix = pre; }
      if (tid == 0 && tau >= Q) sc->sel_bin = -1;
    }
    __syncthreads();
    pc.mark(11);
    const int tb = sc->sel_bin;
    const u64 pref = sc->sel_prefix;
    if (tb < 0) { __syncthreads(); if (last) next_row_copy(); return false; }
    // gather position tb of the order: packed e with bits >> SH == top_v - tb
    {
      uint32_t first = (top_v - (uint32_t)tb) << SH, width = 1u << SH;
      int W4o = W4;
      asm volatile("" : "+r"(W4o), "+r"(first), "+r"(width));
#pragma unroll 1
      for (int c = tid; c < W4o; c += LT) {
        const uint4 u = w4[c];
        if (((u.x - first) < width) | ((u.y - first) < width) | ((u.z - first) < width) | ((u.w - first) < width)) {
          const uint32_t bt[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((bt[j] - first) < width) {
              const int id = 4 * c - mis + j;
              const double ed = l_unpack_e(bt[j]);
              uint32_t q = (uint32_t)__double2loint(__fma_rn(ed, C_lo, magic));
              if (q != (uint32_t)__double2loint(__fma_rn(ed, C_hi, magic))) q = exact_mass(id, C);
              const int s = atomicAdd(&sc->c_n, 1);
              if (s < L_C_CAP) { clist[s].ebits = bt[j]; clist[s].id = id; clist[s].w = q; clist[s].key = 0.f; }
            }
        }
      }
      if (tid < nband && band_kept(tid)) {
        const uint32_t eb = l_pack_e(band[tid].e);
        if ((eb - first) < width) {
          const int s = atomicAdd(&sc->c_n, 1);
          if (s < L_C_CAP) {
            clist[s].ebits = eb; clist[s].id = band[tid].id;
            clist[s].w = (uint32_t)__double2ll_rn(band[tid].e * C); clist[s].key = 0.f;
          }
        }
      }
    }
    pc.mark(12);
    __syncthreads();
    pc.mark(13);
    if (last) next_row_copy();
    const int n = sc->c_n;
    if (n > L_C_CAP) { overflow = true; return false; }
    // coder order: larger e first; equal truncated e (rare with 24 mantissa bits): larger original logit first, fetched
    // on demand; equal logits: lower id first.  Four lanes share one entry, each scans every fourth other entry.
    {
      const int c = tid >> 2, sub = tid & 3;
      const bool live = c < n;
      LCand me = {0u, 0, 0u, 0.f};
      if (live) me = clist[c];
      u64 before = 0;
      if (live) {
        float mkey = 0.f; bool have = false;
        for (int o = sub; o < n; o += 4) {
          const LCand ot = clist[o];
          bool b4;
          if (ot.ebits != me.ebits) b4 = ot.ebits > me.ebits;
          else if (o == c) b4 = false;
          else {
            if (!have) { mkey = g[me.id] + 0.0f; have = true; }
            const float okey = g[ot.id] + 0.0f;
            b4 = (okey != mkey) ? (okey > mkey) : (ot.id < me.id);
          }
          if (b4) before += ot.w;
        }
      }
      before += __shfl_xor_sync(0xffffffffu, before, 1);
      before += __shfl_xor_sync(0xffffffffu, before, 2);
      before += pref;
      if (live && sub == 0 && me.w != 0u && before <= tau && tau < before + me.w) {
        sc->res_idx = me.id; sc->res_before = before; sc->res_w = me.w; sc->res_ebits = me.ebits;
        sc->res_found = 1;
      }
    }
    pc.mark(14);
    __syncthreads();
    return sc->res_found != 0;
  };

  // ------------------------------------------------------------------ overfill (:153-158)
  u64 slack;
  bool truncated = false;
  LCand trunc_e = {0u, 0, 0u, 0.f};
  if (Q > R) {
    if (select_tau(R, false)) {
      truncated = true;
      trunc_e.ebits = sc->res_ebits; trunc_e.id = sc->res_idx; trunc_e.key = g[trunc_e.id] + 0.0f;
      slack = R - sc->res_before;
    } else slack = 0;
    __syncthreads();
    if (tid == 0) { sc->c_n = 0; sc->res_found = 0; }
    __syncthreads();
  } else {
    slack = R - Q;
  }
  pc.mark(7);
  if (overflow) { give_up(L_WHY_BUCKET); return; }

  if (MODE == MODE_ENC) {
    const u64 lo = mp->lo;
    const u64 m_rel = mp->window - lo;                       // next `precision` message bits (:168-171)
    // rank 0 absorbs the slack (:158): bins are [0, q0 + slack), [A_j + slack, A_j + q_j + slack)
    const u64 tau = m_rel >= slack ? m_rel - slack : 0ull;
    int token; u64 nb, nt;
    if (select_tau(tau, true)) {
      token = sc->res_idx;                                   // :172
      const u64 bs = sc->res_before, ws = sc->res_w;
      if (bs == 0) { nb = lo; nt = lo + ws + slack; }
      else { nb = lo + bs + slack; nt = nb + ws; }           // :175-176
    } else {
      // no kept bin holds the target (a gathered bucket did not fit, or the selection ran past the last bin)
      if (tid == 0) l_hand_over(P, slow_ws, row, L_WHY_BUCKET);
      return;
    }
    pc.mark(8);
    if (tid == 0) finish_encode(P, row, slot, token, nb, nt, cand, Q, mp->cursor, mp->mlen);
    pc.mark(9);
  } else {
    // decode after an overfill: the observed token is in range iff it is kept and sorts before the truncation point
    __syncthreads();
    next_row_copy();
    bool in_range = dec_in_range;
    bool flag_ok = dec_flag_ok;
    if (in_range && truncated) {
      const float tkey = g[tok] + 0.0f;
      const bool before_trunc = (tbits != trunc_e.ebits) ? (tbits > trunc_e.ebits)
                              : (tkey != trunc_e.key) ? (tkey > trunc_e.key) : (tok < trunc_e.id);
      if (!before_trunc) { in_range = false; flag_ok = false; }
    }
    const u64 lo = mp->lo;
    const u64 top_mass = (u64)__double2ll_rn(C);             // e of the row maximum is exactly 1
    u64 nb, nt;
    if (!in_range || dec_before == 0) { nb = lo; nt = lo + (in_range ? dec_w : top_mass) + slack; }
    else { nb = lo + dec_before + slack; nt = nb + dec_w; }
    pc.mark(8);
    if (tid == 0) {
      if (P.ntok_total) l_finish_decode(P, row, slot, flag_ok, nb, nt, cand, Q, mp->mlen, mp->olen, mp->oword);
      else finish_decode(P, row, slot, flag_ok, nb, nt, cand, Q);
    }
    pc.mark(9);
  }
}

template <bool UNIT_TEMP, int MODE, bool PROF>
__global__ void __launch_bounds__(LT, 1) ac_lean_kernel(const __grid_constant__ ns_ac_params P, int32_t* slow_ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LSmem sm;
  {
    const uint32_t sbase = l_saddr(smem_raw);
    uint32_t blk = (sbase + 8191u) & ~8191u;                 // histogram at an 8 KB boundary of the shared window
    asm volatile("" : "+r"(blk));
    sm.blk = blk;
  }
  LScal* sc = reinterpret_cast<LScal*>(sm.at(L_OFF_SCAL));
  uint32_t* words = reinterpret_cast<uint32_t*>(sm.at(L_OFF_ROW));   // element id lives at words[id + mis]
  const int tid = threadIdx.x;
  constexpr int HELPER = LT - 32;                          // lane that fetches the next row's scalars
  for (int i = tid; i < NS_EXP_N; i += LT) {               // 2^(i/512) with (i << 11) taken off the high word (l_exp64)
    const double T = c_exp_tab[i];
    reinterpret_cast<uint2*>(sm.at(L_OFF_TAB))[i] = make_uint2((uint32_t)__double2loint(T), (uint32_t)__double2hiint(T) - ((uint32_t)i << 11));
  }
  if (tid == 0) {
    l_mbar_init(&sc->bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sc->issued_row = -1;
  }
  __syncthreads();
  if ((int)blockIdx.x < P.B) {
    if (tid == 0) l_issue_row(P, blockIdx.x, sc, words);
    if (tid >= 32 && tid < 40) l_edges(P, blockIdx.x, words, tid - 32);
    if (tid == HELPER) sc->meta[0] = l_load_meta(P, blockIdx.x, MODE);
  }
  uint32_t parity = 0;
  LClock<PROF> pc;
  pc.on = PROF && (P.prof != nullptr) && tid == 0;
  pc.last = 0;
  for (int k = 0; k < 16; ++k) pc.acc[k] = 0;
  int it = 0;
  for (int row = blockIdx.x; row < P.B; row += gridDim.x, ++it) {
    pc.start();
    __syncthreads();                                       // previous row is finished with shared memory; edges + meta visible
    pc.mark(10);
    LMeta next;
    const int nrow = row + gridDim.x;
    const bool fetch = (tid == HELPER) && (nrow < P.B);
    if (fetch) next = l_load_meta(P, nrow, MODE);          // loads in flight while the row is processed
    lean_row<UNIT_TEMP, MODE, PROF>(P, slow_ws, row, &sc->meta[it & 1], sm, parity, pc);
    if (fetch) sc->meta[(it + 1) & 1] = next;
    if (PROF && pc.on) pc.acc[15] += 1;
  }
  if (PROF && pc.on) for (int k = 0; k < 16; ++k) atomicAdd((unsigned long long*)&P.prof[k], (unsigned long long)pc.acc[k]);
}
