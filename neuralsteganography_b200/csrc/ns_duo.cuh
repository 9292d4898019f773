// ns_duo.cuh -- arithmetic-coder step, threshold form of the cutoff, TWO rows in flight per SM (sm_100a).
// Included by ns_coder.cu after the shared definitions (u64, pack_of, finish_*), inside namespace nsd.
//
// Why: with one row per SM (ns_fast.cuh) the SM idles through every serial stretch of the row (reductions,
// constants, fix-ups, bucket scan, the bulk copy's latency) and through every CTA barrier -- 46 % of the issue
// slots went unused.  A second row does not fit shared memory (201 KB per row), but it fits TENSOR MEMORY: the
// 256 KB of TMEM are, for a kernel without MMAs, a register-file extension -- tcgen05.st / tcgen05.ld 32x32b.x4
// give every thread private, dynamically indexed 16-byte slots (its lane, 256 columns), measured as fast as
// LDS.128 / STS.128 for a thread-private sweep (scripts/microbench_tmem.cu).
//
// One 512-thread CTA per SM = two groups of 8 warps, each with its own row, scratch and named barrier:
//   group 0: row in shared memory, pulled by the bulk-copy engine (cp.async.bulk + mbarrier) one row ahead
//   group 1: row in tensor memory: chunk c = t + 256 j of the row lives in thread t's columns 4j..4j+3; the L pass
//            streams the row from global memory (L2-prefetched one row ahead) and stashes it as it goes
// The groups take rows from a shared counter and never synchronise with each other, so the barrier stalls and serial
// stretches of one row are filled by the sweeps of the other.  Per row (code_base/arithmetic.py:127-190), same
// arithmetic and the same integers as ns_fast.cuh / the exact kernel:
//   L   fp32 online softmax estimate: row max (lowest id), sum of exp, lowest key
//   P1  ONE fp64 exp per element: exact sum of all e_i (fixed order), exact sum of the provisionally cut ones,
//       elements within 2^-10 of the provisional cutoff to a small exact list; the word is replaced by a 32-bit
//       truncation of e_i (0 = not kept)
//   FIX exact normaliser, provisional cutoff verified, band classified; C = range / S_kept
//   P2  q_i = rint(e_i C) from the truncated e_i with an interval test (2 DFMA), undecidable ones redone exactly;
//       encode: integer mass histogram (2048 monotone buckets); decode: conditional sum, no histogram
//   SEL bucket scan -> gather the target bucket -> exact order -> mass before the selected / observed token
//   UPD shared-prefix bits, interval rescale, token / bits out
// Anything unusual queues the row in slow_ws for the exact kernel (same integers).

constexpr int GT = 256;                // threads per group
constexpr int GW = GT / 32;            // warps per group
constexpr int D_NB = 2048;             // histogram buckets
constexpr int D_BPT = D_NB / GT;       // buckets per thread in the scan
constexpr int D_BAND_CAP = 128;
constexpr int D_U_CAP = 128;
constexpr int D_C_CAP = 256;           // gathered entries (live in the histogram words once the scan is in registers)
constexpr int D_MIN_VOCAB = 1024;
constexpr float D_BAND_EPS = 0.0009765625f;
constexpr uint32_t D_TOP = 0xFF000000u;          // packed e of the row maximum (e == 1.0)
constexpr int D_LU = 5;                          // chunks per estimate batch (loads in flight per thread)
constexpr int D_GROUP_BYTES = D_NB * 4 + D_BAND_CAP * 16 + D_U_CAP * 8 + 1024;
constexpr int D_CTA_BYTES = 64;
constexpr int D_FIXED = NS_EXP_N * 8 + D_CTA_BYTES + 2 * D_GROUP_BYTES;
constexpr int D_MAX_VOCAB = (SMEM_LIMIT - D_FIXED) / 4 - 8;
static_assert(D_C_CAP * 16 <= D_NB * 4, "gathered entries alias the histogram");
static_assert(D_MAX_VOCAB >= 50257, "the headline vocabulary must fit");

enum { D_WHY_EST = 1, D_WHY_BAND = 2, D_WHY_VERIFY = 3, D_WHY_RANK = 4, D_WHY_ULIST = 5, D_WHY_BUCKET = 6 };
enum { STORE_SMEM = 0, STORE_TMEM = 1 };

struct DBand { int id; int kept; double e; };
struct DCand { uint32_t ebits; int id; uint32_t w; float key; };
struct DUnd { int id; uint32_t bits; };

// per-row scalars and the few words of the row that plain loads must fetch, all loaded one row ahead by a helper lane
struct __align__(16) DMeta {
  float edge[8];                                 // first and last 16-byte chunk of the row, -inf outside the row
  u64 lo, hi, window;
  int slot, cursor, mlen, tok;
  int phase, olen; uint32_t oword; float xtok;   // decode: output position, partly filled word, logit of the observed token
  float xmask[2];                                // logits of the forbidden tokens (-inf when unused)
  int row, pad;
};

struct DScal {
  DMeta meta[2];
  u64 red[3 * GW];
  u64 bar;                           // mbarrier of the row copy (group 0)
  int cur_it, nxt_it;                // this group's row and the one it will take next (iteration index of the CTA)
  int band_n, u_n, c_n, bail;
  int issued_row, band_kept_n;
  u64 band_cut_int;
  int sh, sel_bin; u64 sel_prefix;
  int res_idx, res_found; u64 res_before, res_w; uint32_t res_ebits; int band_E;
};
static_assert(sizeof(DScal) <= 1024, "DScal too large");

struct DCta { uint32_t tmem_base; int next_it; };

struct DGroup {
  double* tab; uint32_t* hist; DBand* band; DUnd* ulist; DCand* clist; DScal* sc; float* words;
  uint32_t tmem;                     // this thread's TMEM base address (lane quadrant, column half)
  int bar_id;
};

__device__ __forceinline__ void gsync(int id) { asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(GT) : "memory"); }
__device__ __forceinline__ float d_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t d_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void d_mbar_init(u64* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(d_saddr(bar)), "r"(count));
}
__device__ __forceinline__ void d_mbar_expect_tx(u64* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(d_saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void d_bulk_g2s(void* dst, const void* src, uint32_t bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(d_saddr(dst)), "l"(src), "r"(bytes), "r"(d_saddr(bar)) : "memory");
}
__device__ __forceinline__ void d_mbar_wait(u64* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" :: "r"(d_saddr(bar)), "r"(parity) : "memory");
}
// tensor memory as thread-private storage: one float4 per thread per instruction (32 lanes x 4 columns)
__device__ __forceinline__ void tm_st4(uint32_t taddr, const float4 v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               :: "r"(taddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)) : "memory");
}
__device__ __forceinline__ float4 tm_ld4(uint32_t taddr) {
  uint32_t a, b, c, d;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
  return make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d));
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32-bit truncation of e in (2^-255, 1]: bits 59..28 of the double (8 exponent + 24 mantissa bits), monotone,
// unpacked with two shifts; 0 = not kept
__device__ __forceinline__ float d_pack_e(double e) {
  return __uint_as_float(__funnelshift_l((uint32_t)__double2loint(e), (uint32_t)__double2hiint(e), 4));
}
__device__ __forceinline__ double d_unpack_e(float w) {
  const uint32_t b = __float_as_uint(w);
  return __hiloint2double((int)__funnelshift_r(b, 0x3u, 4), (int)(b << 28));
}
__device__ __forceinline__ void d_hist_add(uint32_t* hist, uint32_t bin, uint32_t q) {
  atomicAdd(hist + (q ? bin : (threadIdx.x & (D_NB - 1))), q);
}
__device__ __forceinline__ void d_hand_over(const ns_ac_params& P, int32_t* slow_ws, int row, int why) {
  const int s = atomicAdd(&slow_ws[0], 1);
  slow_ws[2 + s] = row;
  if (P.status) atomicOr(&P.status[row], NS_ST_EST_RETRY | (why << 8));
}

// group-wide sums: two fp64 and one integer with one pair of barriers; fixed order -> deterministic bits
__device__ __forceinline__ void d_sum_ddu(double& a, double& b, u64& c, u64* scratch, int gt, int bid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = a + __shfl_xor_sync(0xffffffffu, a, o);
    b = b + __shfl_xor_sync(0xffffffffu, b, o);
    c = c + __shfl_xor_sync(0xffffffffu, c, o);
  }
  gsync(bid);
  if ((gt & 31) == 0) {
    const int w = gt >> 5;
    scratch[w] = (u64)__double_as_longlong(a);
    scratch[GW + w] = (u64)__double_as_longlong(b);
    scratch[2 * GW + w] = c;
  }
  gsync(bid);
  double ra = __longlong_as_double((long long)scratch[0]);
  double rb = __longlong_as_double((long long)scratch[GW]);
  u64 rc = scratch[2 * GW];
#pragma unroll
  for (int w = 1; w < GW; ++w) {
    ra = ra + __longlong_as_double((long long)scratch[w]);
    rb = rb + __longlong_as_double((long long)scratch[GW + w]);
    rc = rc + scratch[2 * GW + w];
  }
  a = ra; b = rb; c = rc;
}
template <class Op>
__device__ __forceinline__ u64 d_reduce_u(u64 v, Op op, u64* scratch, int gt, int bid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  gsync(bid);
  if ((gt & 31) == 0) scratch[gt >> 5] = v;
  gsync(bid);
  u64 r = scratch[0];
#pragma unroll
  for (int w = 1; w < GW; ++w) r = op(r, scratch[w]);
  return r;
}
__device__ __forceinline__ float d_sum_f(float v, u64* scratch, int gt, int bid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
  gsync(bid);
  if ((gt & 31) == 0) scratch[gt >> 5] = (u64)__float_as_uint(v);
  gsync(bid);
  float r = __uint_as_float((uint32_t)scratch[0]);
#pragma unroll
  for (int w = 1; w < GW; ++w) r = r + __uint_as_float((uint32_t)scratch[w]);
  return r;
}

// bulk copy of the interior chunks 1 .. W4-2 of row `row` into the shared-memory row (one thread)
__device__ __forceinline__ void d_issue_row(const ns_ac_params& P, int row, u64* bar, float* words, int* issued_row) {
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  const int NI = ((mis + P.V + 3) >> 2) - 2;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  d_mbar_expect_tx(bar, (uint32_t)NI * 16u);
  d_bulk_g2s(reinterpret_cast<char*>(words) + 16, reinterpret_cast<const char*>(g - mis) + 16, (uint32_t)NI * 16u, bar);
  *issued_row = row;
}

__device__ __forceinline__ DMeta d_load_meta(const ns_ac_params& P, int row, int mode) {
  DMeta m;
  m.row = row; m.pad = 0;
  m.phase = P.phase ? (int)P.phase[row] : NS_PHASE_CODING;
  m.slot = P.ntok ? P.ntok[row] : 0;
  m.lo = P.lo[row]; m.hi = P.hi[row];
  m.cursor = 0; m.mlen = 0; m.window = 0; m.tok = -1; m.olen = 0; m.oword = 0; m.xtok = -INFINITY;
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
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
    if (m.tok >= 0 && m.tok < P.V) m.xtok = g[m.tok];
  }
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  const int W4 = (mis + P.V + 3) >> 2;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int b = 4 * (k < 4 ? 0 : W4 - 1) - mis + (k & 3);
    m.edge[k] = (b >= 0 && b < P.V) ? g[b] : -INFINITY;
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int id = P.mask_id[k];
    m.xmask[k] = (id >= 0 && id < P.V) ? g[id] : -INFINITY;
  }
  return m;
}

// decode epilogue with the stream's scalars and its partly filled output word already in registers (stores only)
__device__ __forceinline__ void d_finish_decode(const ns_ac_params& P, int row, int slot, bool in_range, u64 nb, u64 nt,
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

struct DClock {
  bool on; long long last; u64 acc[16];
  __device__ __forceinline__ void start() { if (on) last = clock64(); }
  __device__ __forceinline__ void mark(int k) { if (on) { const long long t = clock64(); acc[k] += (u64)(t - last); last = t; } }
};

template <bool UNIT_TEMP, int MODE, int STORE>
__device__ __forceinline__ void duo_row(const ns_ac_params& P, int32_t* slow_ws, const int row, const int nrow, const DMeta* mp,
                                        const DGroup G, const int gt, uint32_t& parity, DClock& pc,
                                        const bool fetch, DMeta& next, bool& have_next) {
  double* tab = G.tab; uint32_t* hist = G.hist; DBand* band = G.band; DUnd* ulist = G.ulist; DCand* clist = G.clist;
  DScal* sc = G.sc; float* words = G.words;
  float4* w4 = reinterpret_cast<float4*>(words);
  const int bid = G.bar_id;
  const int lane = gt & 31, warp = gt >> 5;
  const int V = P.V;
  const double temp = P.temp;
  const float c2 = (float)(1.4426950408889634 / temp);
  const double magic = 6755399441055744.0;
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  const int W4 = (mis + V + 3) >> 2;
  const int per = (W4 + GT - 1) / GT;                      // chunks per thread (warp-uniform loops: tcgen05 ld/st are .aligned)
  const float4 ninf4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // chunk j of this thread = chunk gt + j * GT of the row
  auto ld4 = [&](int j, const float4 neutral) -> float4 {
    if (STORE == STORE_TMEM) return tm_ld4(G.tmem + 4u * (uint32_t)j);
    const int c = gt + j * GT;
    return c < W4 ? w4[c] : neutral;
  };
  auto ld_wait = [&]() { if (STORE == STORE_TMEM) tm_wait_ld(); };
  auto st4 = [&](int j, const float4 v) {
    if (STORE == STORE_TMEM) tm_st4(G.tmem + 4u * (uint32_t)j, v);
    else { const int c = gt + j * GT; if (c < W4) w4[c] = v; }
  };
  auto st_wait = [&]() { if (STORE == STORE_TMEM) tm_wait_st(); };
  auto next_row_copy = [&]() {       // right after a group barrier that follows the row's last read of the shared row
    if (STORE == STORE_SMEM && gt == 0 && nrow < P.B) d_issue_row(P, nrow, &sc->bar, words, &sc->issued_row);
  };
  auto drain = [&]() {               // a skipped row still consumes its copy if the previous row already started it
    if (STORE == STORE_SMEM && sc->issued_row == row) { d_mbar_wait(&sc->bar, parity & 1u); parity ^= 1u; }
  };

  int phase = mp->phase;
  if (phase == NS_PHASE_DONE) { drain(); return; }
  if (MODE != MODE_ENC) phase = NS_PHASE_CODING;
  const int slot = mp->slot;
  if (MODE == MODE_ENC && P.ntok && slot >= P.token_cap) {
    if (gt == 0) {
      if (P.phase) P.phase[row] = NS_PHASE_DONE;
      if (P.status) atomicOr(&P.status[row], NS_ST_TOKEN_OVERFLOW);
    }
    drain();
    return;
  }
  if (MODE == MODE_DEC && P.ntok_total && slot >= mp->mlen) {
    if (gt == 0 && P.phase) P.phase[row] = NS_PHASE_DONE;
    drain();
    return;
  }

  // ------------------------------------------------------------------ L: the row arrives, fp32 estimate
  if (gt == 0) {
    if (STORE == STORE_SMEM && sc->issued_row != row) d_issue_row(P, row, &sc->bar, words, &sc->issued_row);
    sc->band_n = 0; sc->u_n = 0; sc->c_n = 0; sc->bail = 0; sc->band_cut_int = 0; sc->band_kept_n = 0;
  }
  if (STORE == STORE_SMEM && gt < 8) words[4 * (gt < 4 ? 0 : W4 - 1) + (gt & 3)] = mp->edge[gt];
  for (int i = gt; i < D_NB / 4; i += GT) reinterpret_cast<uint4*>(hist)[i] = make_uint4(0, 0, 0, 0);
  if (nrow < P.B && lane == 0) {     // the group's next row into L2, one bulk prefetch per warp
    const char* np = reinterpret_cast<const char*>(P.logits + (size_t)nrow * (size_t)P.ld);
    const char* a0 = reinterpret_cast<const char*>(((uintptr_t)np + 15u) & ~(uintptr_t)15u);
    const int nbytes = (int)(np + (size_t)V * 4 - a0) & ~15;
    const int pw = ((nbytes / GW) + 15) & ~15;
    const int o = warp * pw;
    int n = nbytes - o;
    if (n > pw) n = pw;
    if (n > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0 + o), "r"(n) : "memory");
  }
  pc.mark(0);
  float tm = -3.0e38f, ts = 0.f, ntc = 3.0e38f * c2;       // running max, sum relative to it, -tm * c2
  float kmin = 3.0e38f;
  int ti = 0;
  const float4* g4 = reinterpret_cast<const float4*>(g - mis);   // 16-byte aligned view of the row
  if (STORE == STORE_SMEM) { gsync(bid); d_mbar_wait(&sc->bar, parity & 1u); parity ^= 1u; }   // edges visible, copy landed
  {
    const int nb = (per + D_LU - 1) / D_LU;
    // batch b0 of D_LU chunks; the tensor-memory group reads global memory (L2: prefetched a row ahead) one batch ahead
    auto load_batch = [&](int b0, float4* v) {
#pragma unroll
      for (int u = 0; u < D_LU; ++u) {
        const int j = b0 * D_LU + u, c = gt + j * GT;
        if (STORE == STORE_SMEM) v[u] = (j < per && c < W4) ? w4[c] : ninf4;
        else {
          if (j < per && c > 0 && c < W4 - 1) {
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(g4 + c));
          } else if (c == 0) v[u] = *reinterpret_cast<const float4*>(&mp->edge[0]);
          else if (j < per && c == W4 - 1) v[u] = *reinterpret_cast<const float4*>(&mp->edge[4]);
          else v[u] = ninf4;
        }
      }
    };
    float4 nx[D_LU];
    if (STORE == STORE_TMEM) load_batch(0, nx);
#pragma unroll 1
    for (int b0 = 0; b0 < nb; ++b0) {
      float4 v[D_LU];
      if (STORE == STORE_TMEM) {
#pragma unroll
        for (int u = 0; u < D_LU; ++u) v[u] = nx[u];
        if (b0 + 1 < nb) load_batch(b0 + 1, nx);
      } else load_batch(b0, v);
      float cm = -INFINITY;
#pragma unroll
      for (int u = 0; u < D_LU; ++u) cm = fmaxf(cm, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
      if (cm > tm) {                                       // rare after the first batches
        ts *= d_ex2((tm - cm) * c2);
        tm = cm;
        ntc = -cm * c2;
        bool found = false;
#pragma unroll
        for (int u = 0; u < D_LU; ++u) {
          const int bb = 4 * (gt + (b0 * D_LU + u) * GT) - mis;
          const float xs[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) if (!found && xs[e] == cm) { ti = bb + e; found = true; }
        }
      }
      float part = 0.f;
#pragma unroll
      for (int u = 0; u < D_LU; ++u) {
        part += (d_ex2(fmaf(v[u].x, c2, ntc)) + d_ex2(fmaf(v[u].y, c2, ntc))) + (d_ex2(fmaf(v[u].z, c2, ntc)) + d_ex2(fmaf(v[u].w, c2, ntc)));
        const int c = gt + (b0 * D_LU + u) * GT;
        const float lo4 = fminf(fminf(v[u].x, v[u].y), fminf(v[u].z, v[u].w));
        if (c > 0 && c < W4 - 1) kmin = fminf(kmin, lo4);
        else {                                             // edge chunks carry -inf padding: not part of the key range
          const float xs[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) if (xs[e] > -INFINITY) kmin = fminf(kmin, xs[e]);
        }
        if (STORE == STORE_TMEM && b0 * D_LU + u < per) tm_st4(G.tmem + 4u * (uint32_t)(b0 * D_LU + u), v[u]);
      }
      ts += part;
    }
    st_wait();
  }
  pc.mark(1);
  // row reductions with one barrier: warp partials (max key with lowest id, sum rescaled to the warp's max, lowest key)
  // meet in shared memory, every warp combines them the same way (identical bits in every thread)
  float M, ssum, key_min;
  int top_id;
  {
    const uint32_t ok = ns_f32_orderable(tm + 0.0f);
    const uint32_t wk = __reduce_max_sync(0xffffffffu, ok);
    const int wi = __reduce_min_sync(0xffffffffu, ok == wk ? ti : 0x7fffffff);
    const float wm = key_of_pack((u64)wk << 32);
    float wts = ts * d_ex2((tm - wm) * c2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wts += __shfl_xor_sync(0xffffffffu, wts, o);
    const uint32_t wmin = __reduce_min_sync(0xffffffffu, ns_f32_orderable(kmin));
    uint4* red4 = reinterpret_cast<uint4*>(sc->red);
    if (lane == 0) red4[warp] = make_uint4(wk, (uint32_t)wi, __float_as_uint(wts), wmin);
    gsync(bid);
    const uint4 pr = red4[lane < GW ? lane : 0];
    const uint32_t mk = __reduce_max_sync(0xffffffffu, pr.x);
    top_id = __reduce_min_sync(0xffffffffu, pr.x == mk ? (int)pr.y : 0x7fffffff);
    M = key_of_pack((u64)mk << 32);
    key_min = key_of_pack((u64)__reduce_min_sync(0xffffffffu, pr.w) << 32);
    float part = lane < GW ? __uint_as_float(pr.z) * d_ex2((key_of_pack((u64)pr.x << 32) - M) * c2) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    ssum = part;
    // forbidden tokens (code_base/arithmetic.py:124-125): probability exactly 0 -- out of the estimate, -inf in the stored row
    bool remax = false;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int id = P.mask_id[k];
      if (id >= 0 && id < V) {
        { const float xm = mp->xmask[k]; if (xm > -INFINITY) ssum -= d_ex2((xm - M) * c2); }
        if (id == top_id) remax = true;
        const int ce = id + mis, c = ce >> 2, ow = (c % GT) >> 5, jo = c / GT;
        if (STORE == STORE_SMEM) { if (gt == k) words[ce] = -INFINITY; }
        else if (warp == ow) {                             // the owner warp rewrites its chunk (warp-uniform: .aligned)
          float4 v = tm_ld4(G.tmem + 4u * (uint32_t)jo);
          tm_wait_ld();
          if (lane == ((c % GT) & 31)) reinterpret_cast<float*>(&v)[ce & 3] = -INFINITY;
          tm_st4(G.tmem + 4u * (uint32_t)jo, v);
          tm_wait_st();
        }
      }
    }
    gsync(bid);                                            // the -inf is visible to whoever sweeps the row next
    if (remax) {                                           // rare: the row maximum itself was forbidden
      u64 pm = 0;
      for (int j = 0; j < per; ++j) {
        const float4 v = ld4(j, ninf4);
        ld_wait();
        const int b = 4 * (gt + j * GT) - mis;
        u64 p;
        p = pack_of(v.x + 0.0f, b); if (v.x > -INFINITY) pm = p > pm ? p : pm;
        p = pack_of(v.y + 0.0f, b + 1); if (v.y > -INFINITY) pm = p > pm ? p : pm;
        p = pack_of(v.z + 0.0f, b + 2); if (v.z > -INFINITY) pm = p > pm ? p : pm;
        p = pack_of(v.w + 0.0f, b + 3); if (v.w > -INFINITY) pm = p > pm ? p : pm;
      }
      pm = d_reduce_u(pm, OpMaxU(), sc->red, gt, bid);
      M = key_of_pack(pm);
      top_id = id_of_pack(pm);
      float s = 0.f;
      for (int j = 0; j < per; ++j) {
        const float4 v = ld4(j, ninf4);
        ld_wait();
        s += d_ex2((v.x - M) * c2) + d_ex2((v.y - M) * c2) + d_ex2((v.z - M) * c2) + d_ex2((v.w - M) * c2);
      }
      ssum = d_sum_f(s, sc->red, gt, bid);
      gsync(bid);
    }
  }

  if (MODE == MODE_ENC && phase == NS_PHASE_TAIL) {
    next_row_copy();
    if (gt == 0) finish_tail(P, row, slot, top_id);
    return;
  }

  // ------------------------------------------------------------------ row constants (every thread, same bits)
  const u64 lo = mp->lo, hi = mp->hi;
  const u64 R = hi - lo;                                   // arithmetic.py:140
  const double thr = __ddiv_rn(1.0, (double)R);            // :141
  const double Md = (double)M;
  const double dm = UNIT_TEMP ? Md : __ddiv_rn(Md, temp);
  // provisional cutoff from the estimate: p >= 1/R  <=>  key >= M + temp * ln(sum / R); the band absorbs its error and
  // the split is verified exactly after the exp pass
  const double theta_est = thr * (double)ssum;
  const float tf = (float)temp;
  const float key_th = fmaf(tf * 0.6931471805599453f, __log2f((float)theta_est), M);
  const float kappa_hi = key_th + tf * D_BAND_EPS, kappa_lo = key_th - tf * D_BAND_EPS;
  const float clamp_key = (float)(Md - 700.0 * temp);
  const int band_E = ((__double2hiint(theta_est) >> 20) & 0x7ff) - 1024;   // ilogb(theta_est) - 1
  if (!(ssum > 0.0f) || !(R >= 2) || !(kappa_lo > clamp_key)) {
    next_row_copy();
    if (gt == 0) d_hand_over(P, slow_ws, row, D_WHY_EST);
    return;
  }
  auto a_of = [&](float key) -> double {                   // (double(x)/temp) - (double(max)/temp), :128-130
    double x = (double)fmaxf(key, clamp_key);
    if (!UNIT_TEMP) x = __ddiv_rn(x, temp);
    return x - dm;
  };
  auto band_push = [&](int id, double e) {
    const int s = atomicAdd(&sc->band_n, 1);
    if (s < D_BAND_CAP) { band[s].id = id; band[s].kept = 0; band[s].e = e; }
  };
  pc.mark(2);
  // ------------------------------------------------------------------ P1: the fp64 exp pass
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
  double accl0 = 0.0, accl1 = 0.0, accl2 = 0.0, accl3 = 0.0;
  int cnt_hi = 0;
  const bool need_count = P.topk < V;                      // otherwise only "at least 2 kept" matters
#pragma unroll 1
  for (int j = 0; j < per; ++j) {
    const float4 v = ld4(j, ninf4);
    ld_wait();
    const int b = 4 * (gt + j * GT) - mis;
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
    o.x = h0 ? d_pack_e(e0) : 0.0f;
    o.y = h1 ? d_pack_e(e1) : 0.0f;
    o.z = h2 ? d_pack_e(e2) : 0.0f;
    o.w = h3 ? d_pack_e(e3) : 0.0f;
    st4(j, o);
    if (!((h0 | l0) & (h1 | l1) & (h2 | l2) & (h3 | l3))) {   // rare: inside the guard band
      if (!(h0 | l0)) band_push(b, e0);
      if (!(h1 | l1)) band_push(b + 1, e1);
      if (!(h2 | l2)) band_push(b + 2, e2);
      if (!(h3 | l3)) band_push(b + 3, e3);
    }
  }
  st_wait();
  pc.mark(3);
  // the next row's scalars: loads issued here (not across the exp pass: its registers are all taken), stored after the row
  if (fetch) { next = d_load_meta(P, nrow, MODE); have_next = true; }
  double sum_all = (acc0 + acc1) + (acc2 + acc3);          // softmax normaliser, :130
  double sum_lo = (accl0 + accl1) + (accl2 + accl3);
  u64 n_hi = (u64)cnt_hi;
  d_sum_ddu(sum_all, sum_lo, n_hi, sc->red, gt, bid);
  const double inv = __ddiv_rn(1.0, sum_all);
  const int nband = sc->band_n;
  const double band_scale = scalbn(1.0, 52 - band_E);
  // ------------------------------------------------------------------ FIX: exact classification
  const float kappa_lo_pred = nextafterf(kappa_lo, -INFINITY);
  if (nband <= D_BAND_CAP && gt < nband) {
    const double e = band[gt].e;
    const bool k = (e * inv) >= thr;                       // p_i >= 1/range, :69
    band[gt].kept = k ? 1 : 0;
    if (k) atomicAdd(&sc->band_kept_n, 1);
    else atomicAdd(&sc->band_cut_int, (u64)__double2ull_rz(e * band_scale));   // exact, order-free
  }
  if (gt == GT - 64) {
    // the provisional split is valid iff exp is monotone and both band edges classify as assumed
    const double e_hi = ns_exp64_core(a_of(kappa_hi), tab);
    const double e_lo = ns_exp64_core(a_of(kappa_lo_pred), tab);
    int bail = 0;
    if (nband > D_BAND_CAP) bail = D_WHY_BAND;
    else if (!((e_hi * inv) >= thr) || ((e_lo * inv) >= thr)) bail = D_WHY_VERIFY;
    sc->bail = bail;
  }
  if (gt == GT - 96) {
    // bucket shift: every kept element has e >= e(max(kappa_lo_pred, lowest logit)) > 0
    const float e_min = d_pack_e(ns_exp64_core(a_of(fmaxf(kappa_lo_pred, key_min)), tab));
    const uint32_t span = D_TOP - __float_as_uint(e_min);
    int sh = 0;
    while ((span >> sh) > (uint32_t)(D_NB - 1)) ++sh;
    sc->sh = sh;
  }
  gsync(bid);
  const u64 cand = n_hi + (u64)sc->band_kept_n;            // only counted when topk < V
  const double sum_bc = (double)sc->band_cut_int * scalbn(1.0, band_E - 52);
  const double S = (sum_all - sum_lo) - sum_bc;            // sum of the kept e_i
  const bool form_ok = need_count ? (cand >= 2 && cand <= (u64)P.topk) : ((inv >= thr) && (S > 1.0 + 0.5 * thr));
  if (sc->bail || !form_ok) {                              // rank form (top-k inside the cutoff set) -> exact kernel
    const int why = sc->bail ? sc->bail : D_WHY_RANK;
    gsync(bid);
    next_row_copy();
    if (gt == 0) d_hand_over(P, slow_ws, row, why);
    return;
  }
  const double C = __ddiv_rn((double)R, S);                // :146
  const double C_lo = C * (1.0 - 2.220446049250313e-16);
  const double C_hi = C * (1.0 + 5.960464477539063e-08 + 9.094947017729282e-13);   // e < e_trunc * (1 + 2^-24)
  const int SH = sc->sh;
  auto bin_of_e = [&](float e32) -> uint32_t { return (D_TOP - __float_as_uint(e32)) >> SH; };
  auto exact_mass = [&](int id) -> uint32_t {              // bin width from the original logit (the exact kernel's formula)
    const float key = g[id] + 0.0f;
    return (uint32_t)__double2ll_rn(ns_exp64_core(a_of(key), tab) * C);
  };
  auto quick_mass = [&](float e32, uint32_t* q) -> bool {  // decided from the truncated e alone?
    const double ed = d_unpack_e(e32);
    const uint32_t ql = (uint32_t)ns_double_as_u64(__fma_rn(ed, C_lo, magic));
    const uint32_t qh = (uint32_t)ns_double_as_u64(__fma_rn(ed, C_hi, magic));
    *q = ql;
    return ql == qh;
  };
  // packed e of token `tok` as the exp pass stored it (0 = not in the certain kept set), recomputed from its logit
  auto stored_e32 = [&](int tok, float x) -> float {
    if (tok == P.mask_id[0] || tok == P.mask_id[1]) return 0.0f;
    return (x >= kappa_hi) ? d_pack_e(ns_exp64_core(a_of(x), tab)) : 0.0f;
  };
  auto und_push = [&](int id, float e32) {
    const int s = atomicAdd(&sc->u_n, 1);
    if (s < D_U_CAP) { ulist[s].id = id; ulist[s].bits = __float_as_uint(e32); }
  };
  pc.mark(4);
  // ------------------------------------------------------------------ decode without a histogram
  if (MODE == MODE_DEC) {
    int tok = mp->tok;
    float xt = mp->xtok;
    if (tok < 0 || tok >= V) { tok = top_id; xt = M; }
    float e32t = stored_e32(tok, xt);
    bool tok_in_band = false;
    if (__float_as_uint(e32t) == 0u) {
      for (int k = 0; k < nband; ++k)
        if (band[k].id == tok && band[k].kept) { e32t = d_pack_e(band[k].e); tok_in_band = true; }
    }
    const uint32_t tbits = __float_as_uint(e32t);          // 0: the token is not in the kept set
    uint32_t qs = 0, bs32 = 0;
    auto tie_push = [&](int id, uint32_t q) {
      const int s2 = atomicAdd(&sc->c_n, 1);
      if (s2 < D_C_CAP) { clist[s2].ebits = tbits; clist[s2].id = id; clist[s2].w = q; clist[s2].key = 0.0f; }
    };
    auto d2_slow = [&](const float4 v, const int b, const uint32_t* qv, const bool* kv) {
      const float ev[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t bits = __float_as_uint(ev[e]);
        if (!kv[e]) und_push(b + e, ev[e]);
        else if (qv[e]) {
          qs += qv[e];
          if (bits > tbits) bs32 += qv[e];
          else if (bits == tbits && b + e != tok) tie_push(b + e, qv[e]);
        }
      }
    };
#pragma unroll 1
    for (int j = 0; j < per; j += 2) {
      const float4 v = ld4(j, zero4);
      const float4 w = (j + 1 < per) ? ld4(j + 1, zero4) : zero4;
      ld_wait();
      const int b = 4 * (gt + j * GT) - mis;
      uint32_t q[8];
      bool k[8];
      k[0] = quick_mass(v.x, &q[0]); k[1] = quick_mass(v.y, &q[1]); k[2] = quick_mass(v.z, &q[2]); k[3] = quick_mass(v.w, &q[3]);
      k[4] = quick_mass(w.x, &q[4]); k[5] = quick_mass(w.y, &q[5]); k[6] = quick_mass(w.z, &q[6]); k[7] = quick_mass(w.w, &q[7]);
      const uint32_t bt[8] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w),
                              __float_as_uint(w.x), __float_as_uint(w.y), __float_as_uint(w.z), __float_as_uint(w.w)};
      bool tie = false;
#pragma unroll
      for (int e = 0; e < 8; ++e) tie |= (bt[e] == tbits);
      if ((k[0] & k[1] & k[2] & k[3] & k[4] & k[5] & k[6] & k[7]) && !(tie && tbits != 0u)) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { qs += q[e]; bs32 += bt[e] > tbits ? q[e] : 0u; }
      } else {
        d2_slow(v, b, q, k);
        d2_slow(w, b + 4 * GT, q + 4, k + 4);
      }
    }
    pc.mark(5);
    gsync(bid);
    const int nu = sc->u_n;
    if (nu > D_U_CAP) { gsync(bid); next_row_copy(); if (gt == 0) d_hand_over(P, slow_ws, row, D_WHY_ULIST); return; }
    for (int u = gt; u < nu; u += GT) {                    // widths the truncated e could not decide
      const int id = ulist[u].id;
      const uint32_t m = exact_mass(id), bits = ulist[u].bits;
      qs += m;
      if (bits > tbits) bs32 += m;
      else if (bits == tbits && id != tok && m) tie_push(id, m);
    }
    if (gt < nband && band[gt].kept) {                     // the exact-list band
      const double e = band[gt].e;
      const uint32_t m = (uint32_t)__double2ll_rn(e * C), bits = __float_as_uint(d_pack_e(e));
      qs += m;
      if (bits > tbits) bs32 += m;
      else if (bits == tbits && band[gt].id != tok && m) tie_push(band[gt].id, m);
    }
    gsync(bid);
    const int nt_ties = sc->c_n;
    if (nt_ties > D_C_CAP) { gsync(bid); next_row_copy(); if (gt == 0) d_hand_over(P, slow_ws, row, D_WHY_BUCKET); return; }
    if (gt < nt_ties) {                                    // same truncated e as the token: original logit, then id
      const float key = g[clist[gt].id] + 0.0f, tkey = xt + 0.0f;
      if (key > tkey || (key == tkey && clist[gt].id < tok)) bs32 += clist[gt].w;
    }
    u64 Qd = (u64)qs, Bd = (u64)bs32;
    {                                                      // exact integer sums: any order
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Qd += __shfl_xor_sync(0xffffffffu, Qd, o);
        Bd += __shfl_xor_sync(0xffffffffu, Bd, o);
      }
      gsync(bid);
      if (lane == 0) { sc->red[warp] = Qd; sc->red[GW + warp] = Bd; }
      gsync(bid);
      Qd = 0; Bd = 0;
#pragma unroll
      for (int w = 0; w < GW; ++w) { Qd += sc->red[w]; Bd += sc->red[GW + w]; }
    }
    pc.mark(6);
    if (Qd <= R) {
      next_row_copy();                                     // the row buffer is not read again
      const u64 slack = R - Qd;                            // :158
      const u64 top_mass = (u64)__double2ll_rn(C);         // e of the row maximum is exactly 1
      bool in_range = tbits != 0u;
      u64 ws = top_mass, bsum = 0;
      int token = top_id;
      if (in_range) {
        uint32_t qt = 0;
        if (tok_in_band) { for (int k = 0; k < nband; ++k) if (band[k].id == tok) qt = (uint32_t)__double2ll_rn(band[k].e * C); }
        else if (!quick_mass(e32t, &qt)) qt = exact_mass(tok);
        ws = qt; bsum = Bd; token = tok;
      }
      u64 nb, nt;
      if (token == top_id) { nb = lo; nt = lo + ws + slack; }   // :342 / :347-348
      else { nb = lo + bsum + slack; nt = nb + ws; }
      pc.mark(8);
      if (gt == 0) {
        if (P.ntok_total) d_finish_decode(P, row, slot, in_range, nb, nt, cand, Qd, mp->mlen, mp->olen, mp->oword);
        else finish_decode(P, row, slot, in_range, nb, nt, cand, Qd);
      }
      pc.mark(9);
      return;
    }
    // overfill: the general path needs clean lists and a clean histogram (the tie list lived in its words)
    gsync(bid);
    if (gt == 0) { sc->u_n = 0; sc->c_n = 0; }
    for (int i = gt; i < D_NB / 4; i += GT) reinterpret_cast<uint4*>(hist)[i] = make_uint4(0, 0, 0, 0);
    gsync(bid);
  }
  // ------------------------------------------------------------------ P2: integer bin widths -> mass histogram
  auto p2_slow = [&](const float4 v, const int b, const uint32_t* qv, const bool* kv) {
    const float ev[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (kv[e]) { if (qv[e]) atomicAdd(&hist[bin_of_e(ev[e])], qv[e]); }
      else und_push(b + e, ev[e]);
    }
  };
#pragma unroll 1
  for (int j = 0; j < per; j += 2) {
    const float4 v = ld4(j, zero4);
    const float4 w = (j + 1 < per) ? ld4(j + 1, zero4) : zero4;
    ld_wait();
    const int b = 4 * (gt + j * GT) - mis;
    uint32_t q[8];
    bool k[8];
    k[0] = quick_mass(v.x, &q[0]); k[1] = quick_mass(v.y, &q[1]); k[2] = quick_mass(v.z, &q[2]); k[3] = quick_mass(v.w, &q[3]);
    k[4] = quick_mass(w.x, &q[4]); k[5] = quick_mass(w.y, &q[5]); k[6] = quick_mass(w.z, &q[6]); k[7] = quick_mass(w.w, &q[7]);
    if (k[0] & k[1] & k[2] & k[3] & k[4] & k[5] & k[6] & k[7]) {   // e32 == 0 (not kept) yields q == 0
      d_hist_add(hist, bin_of_e(v.x) & (D_NB - 1), q[0]);
      d_hist_add(hist, bin_of_e(v.y) & (D_NB - 1), q[1]);
      d_hist_add(hist, bin_of_e(v.z) & (D_NB - 1), q[2]);
      d_hist_add(hist, bin_of_e(v.w) & (D_NB - 1), q[3]);
      d_hist_add(hist, bin_of_e(w.x) & (D_NB - 1), q[4]);
      d_hist_add(hist, bin_of_e(w.y) & (D_NB - 1), q[5]);
      d_hist_add(hist, bin_of_e(w.z) & (D_NB - 1), q[6]);
      d_hist_add(hist, bin_of_e(w.w) & (D_NB - 1), q[7]);
    } else {
      p2_slow(v, b, q, k);
      p2_slow(w, b + 4 * GT, q + 4, k + 4);
    }
  }
  pc.mark(5);
  gsync(bid);
  const int nu = sc->u_n;
  if (nu > D_U_CAP) { gsync(bid); next_row_copy(); if (gt == 0) d_hand_over(P, slow_ws, row, D_WHY_ULIST); return; }
  for (int u = gt; u < nu; u += GT) atomicAdd(&hist[bin_of_e(__uint_as_float(ulist[u].bits))], exact_mass(ulist[u].id));
  if (gt < nband && band[gt].kept) {
    const double e = band[gt].e;
    atomicAdd(&hist[bin_of_e(d_pack_e(e))], (uint32_t)__double2ll_rn(e * C));
  }
  gsync(bid);

  // ------------------------------------------------------------------ SEL: bucket scan, kept in registers
  u64 hloc[D_BPT];
  u64 hexcl;                                               // mass in all buckets before this thread's first one
  u64 Q;
  {
    u64 tsum = 0;
#pragma unroll
    for (int b = 0; b < D_BPT; ++b) { hloc[b] = hist[gt * D_BPT + b]; tsum += hloc[b]; }
    u64 inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sc->red[warp] = inc;
    gsync(bid);                                            // also: every thread has its buckets, the words are free
    u64 woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < GW; ++w) { const u64 x = sc->red[w]; if (w < warp) woff += x; tot += x; }
    hexcl = woff + inc - tsum;
    Q = tot;
  }
  pc.mark(6);
  auto locate = [&](u64 tau) {                             // first bucket whose inclusive prefix exceeds tau
    if (gt == 0) { sc->sel_bin = -1; sc->sel_prefix = 0; }
    gsync(bid);
    u64 excl = hexcl;
#pragma unroll
    for (int b = 0; b < D_BPT; ++b) {
      if (hloc[b] != 0 && excl <= tau && tau < excl + hloc[b]) { sc->sel_bin = gt * D_BPT + b; sc->sel_prefix = excl; }
      excl += hloc[b];
    }
    gsync(bid);
  };
  auto prefix_of = [&](int tb) {                           // mass in all buckets before bucket tb
    gsync(bid);
    if (tb >= gt * D_BPT && tb < (gt + 1) * D_BPT) {
      u64 excl = hexcl;
      const int off = tb - gt * D_BPT;
#pragma unroll
      for (int b = 0; b < D_BPT; ++b) if (b < off) excl += hloc[b];
      sc->sel_prefix = excl;
    }
    gsync(bid);
  };
  auto collect = [&](int tb) -> int {                      // gather bucket tb: every kept element in it with its exact mass
    if (gt == 0) sc->c_n = 0;
    gsync(bid);
    auto gather4 = [&](const float4 v, const int j) {
      const float ev[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (bin_of_e(ev[e]) == (uint32_t)tb) {             // packed 0 (not kept) maps far outside the histogram
          const int id = 4 * (gt + j * GT) - mis + e;
          uint32_t q;
          if (!quick_mass(ev[e], &q)) q = exact_mass(id);
          const int s = atomicAdd(&sc->c_n, 1);
          if (s < D_C_CAP) { clist[s].ebits = __float_as_uint(ev[e]); clist[s].id = id; clist[s].w = q; clist[s].key = 0.0f; }
        }
      }
    };
    auto hit4 = [&](const float4 v) -> bool {
      return (bin_of_e(v.x) == (uint32_t)tb) | (bin_of_e(v.y) == (uint32_t)tb) | (bin_of_e(v.z) == (uint32_t)tb) | (bin_of_e(v.w) == (uint32_t)tb);
    };
#pragma unroll 1
    for (int j = 0; j < per; j += 3) {
      const float4 va = ld4(j, zero4);
      const float4 vb = (j + 1 < per) ? ld4(j + 1, zero4) : zero4;
      const float4 vc = (j + 2 < per) ? ld4(j + 2, zero4) : zero4;
      ld_wait();
      const bool ha = hit4(va), hb = hit4(vb), hc = hit4(vc);
      if (ha | hb | hc) {
        if (ha) gather4(va, j);
        if (hb) gather4(vb, j + 1);
        if (hc) gather4(vc, j + 2);
      }
    }
    if (gt < nband && band[gt].kept) {
      const float e32 = d_pack_e(band[gt].e);
      if (bin_of_e(e32) == (uint32_t)tb) {
        const int s = atomicAdd(&sc->c_n, 1);
        if (s < D_C_CAP) {
          clist[s].ebits = __float_as_uint(e32); clist[s].id = band[gt].id;
          clist[s].w = (uint32_t)__double2ll_rn(band[gt].e * C); clist[s].key = 0.0f;
        }
      }
    }
    gsync(bid);
    const int n = sc->c_n;
    return n > D_C_CAP ? -1 : n;                           // dense bucket: the exact kernel redoes the row
  };
  auto cand_before = [&](const DCand& x, const DCand& y) -> bool {   // larger e, then larger logit, then lower id
    if (x.ebits != y.ebits) return x.ebits > y.ebits;
    if (x.key != y.key) return x.key > y.key;
    return x.id < y.id;
  };
  auto resolve = [&](int n, u64 prefix, bool by_token, u64 tau, int want_id) -> bool {
    for (int c = gt; c < n; c += GT) {                     // entries sharing a packed e need the original logit
      const uint32_t eb = clist[c].ebits;
      bool d = false;
      for (int o = 0; o < n; ++o) d |= (o != c) && (clist[o].ebits == eb);
      if (d) clist[c].key = g[clist[c].id] + 0.0f;
    }
    if (gt == 0) sc->res_found = 0;
    gsync(bid);
    for (int c = gt; c < n; c += GT) {
      const DCand me = clist[c];
      u64 before = prefix;
      for (int o = 0; o < n; ++o) {
        const DCand ot = clist[o];
        if (o != c && cand_before(ot, me)) before += ot.w;
      }
      const bool hit = by_token ? (me.id == want_id) : (me.w != 0 && before <= tau && tau < before + me.w);
      if (hit) { sc->res_idx = me.id; sc->res_before = before; sc->res_w = me.w; sc->res_ebits = me.ebits; sc->res_found = 1; }
    }
    gsync(bid);
    return sc->res_found != 0;
  };
  bool overflow = false;
  uint32_t sel_ebits = 0;
  auto select_tau = [&](u64 tau, int* idx, u64* before, u64* w, bool last) -> bool {
    locate(tau);
    const int tb = sc->sel_bin;
    const u64 pref = sc->sel_prefix;
    if (tb < 0) { if (last) next_row_copy(); return false; }
    const int n = collect(tb);
    if (last) next_row_copy();
    if (n < 0) { overflow = true; return false; }
    const bool f = resolve(n, pref, false, tau, 0);
    *idx = sc->res_idx; *before = sc->res_before; *w = sc->res_w; sel_ebits = sc->res_ebits;
    gsync(bid);
    return f;
  };

  // ------------------------------------------------------------------ overfill (:153-158)
  u64 slack;
  bool truncated = false;
  DCand trunc_e = {0u, 0, 0u, 0.0f};
  if (Q > R) {
    int jx; u64 bj, wj;
    if (select_tau(R, &jx, &bj, &wj, false)) {
      truncated = true;
      trunc_e.ebits = sel_ebits; trunc_e.id = jx; trunc_e.key = g[jx] + 0.0f;
      slack = R - bj;
    } else slack = 0;
  } else {
    slack = R - Q;
  }
  pc.mark(7);
  if (overflow) { gsync(bid); next_row_copy(); if (gt == 0) d_hand_over(P, slow_ws, row, D_WHY_BUCKET); return; }
  const u64 top_mass = (u64)__double2ll_rn(C);             // e of the row maximum is exactly 1
  u64 nb, nt;
  if (MODE == MODE_ENC) {
    const u64 m_rel = mp->window - lo;                    // next `precision` message bits (:168-171)
    int token;
    if (m_rel < top_mass + slack) {                        // rank 0 absorbs the slack (:158)
      gsync(bid);
      next_row_copy();
      token = top_id; nb = lo; nt = lo + top_mass + slack;
    } else {
      int s; u64 bs, ws;
      if (!select_tau(m_rel - slack, &s, &bs, &ws, true)) {
        if (overflow) { if (gt == 0) d_hand_over(P, slow_ws, row, D_WHY_BUCKET); return; }
        s = top_id; bs = 0; ws = top_mass;
        if (gt == 0 && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW);
      }
      token = s;                                           // :172
      if (s == top_id) { nb = lo; nt = lo + ws + slack; }
      else { nb = lo + bs + slack; nt = nb + ws; }         // :175-176
    }
    pc.mark(8);
    if (gt == 0) finish_encode(P, row, slot, token, nb, nt, cand, Q, mp->cursor, mp->mlen);
    pc.mark(9);
  } else {
    int tok = mp->tok;
    float xt = mp->xtok;
    if (tok < 0 || tok >= V) { tok = top_id; xt = M; }
    float e32t = stored_e32(tok, xt);
    if (__float_as_uint(e32t) == 0u) {
      for (int k = 0; k < nband; ++k)
        if (band[k].id == tok && band[k].kept) e32t = d_pack_e(band[k].e);
    }
    bool in_range = __float_as_uint(e32t) != 0u;
    u64 bs = 0, ws = top_mass;
    int token = top_id;
    if (in_range) {
      const int tb = (int)bin_of_e(e32t);
      prefix_of(tb);
      const u64 pref = sc->sel_prefix;
      const int n = collect(tb);
      next_row_copy();
      if (n < 0) { if (gt == 0) d_hand_over(P, slow_ws, row, D_WHY_BUCKET); return; }
      if (resolve(n, pref, true, 0, tok)) { bs = sc->res_before; ws = sc->res_w; token = tok; }
      else in_range = false;
      gsync(bid);
      if (in_range && truncated) {
        DCand me = {__float_as_uint(e32t), tok, 0u, xt + 0.0f};
        const DCand tr = trunc_e;
        if (!cand_before(me, tr)) { in_range = false; token = top_id; bs = 0; ws = top_mass; }
      }
    } else {
      gsync(bid);
      next_row_copy();
    }
    if (token == top_id) { nb = lo; nt = lo + ws + slack; }   // :342 / :347-348
    else { nb = lo + bs + slack; nt = nb + ws; }
    pc.mark(8);
    if (gt == 0) finish_decode(P, row, slot, in_range, nb, nt, cand, Q);
    pc.mark(9);
  }
}

template <bool UNIT_TEMP, int MODE, int STORE>
__device__ __forceinline__ void duo_group(const ns_ac_params& P, int32_t* slow_ws, DGroup G, DCta* cta, const int gt) {
  DScal* sc = G.sc;
  const int bid = G.bar_id;
  constexpr int HELPER = GT - 32;                          // lane that fetches the next row's scalars
  const int stride = (int)gridDim.x, base = (int)blockIdx.x;
  if (gt == 0) {
    if (STORE == STORE_SMEM) {
      d_mbar_init(&sc->bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    sc->issued_row = -1;
    sc->cur_it = atomicAdd(&cta->next_it, 1);
    sc->nxt_it = atomicAdd(&cta->next_it, 1);
  }
  gsync(bid);
  {
    const int r0 = base + sc->cur_it * stride;
    if (gt == HELPER && r0 < P.B) sc->meta[0] = d_load_meta(P, r0, MODE);
  }
  uint32_t parity = 0;
  DClock pc;
  pc.on = (P.prof != nullptr) && gt == 0;
  pc.last = 0;
  for (int k = 0; k < 16; ++k) pc.acc[k] = 0;
  for (int it = 0;; ++it) {
    pc.start();
    gsync(bid);                                            // previous row is finished with the group's memory
    pc.mark(10);
    const int row = base + sc->cur_it * stride;
    const int nrow_raw = base + sc->nxt_it * stride;
    if (row >= P.B) break;
    const int nrow = nrow_raw < P.B ? nrow_raw : P.B;
    gsync(bid);                                            // everyone has the row numbers: thread 0 may claim the one after
    int claimed = 0;
    if (gt == 0) claimed = atomicAdd(&cta->next_it, 1);
    DMeta next;
    bool have_next = false;
    const bool fetch = (gt == HELPER) && (nrow < P.B);
    duo_row<UNIT_TEMP, MODE, STORE>(P, slow_ws, row, nrow, &sc->meta[it & 1], G, gt, parity, pc, fetch, next, have_next);
    if (fetch) {
      if (!have_next) next = d_load_meta(P, nrow, MODE);   // the row left early (finished stream, hand-over)
      sc->meta[(it + 1) & 1] = next;
    }
    if (gt == 0) { sc->cur_it = sc->nxt_it; sc->nxt_it = claimed; }
    if (pc.on) pc.acc[15] += 1;
  }
  if (pc.on) for (int k = 0; k < 16; ++k) atomicAdd((unsigned long long*)&P.prof[16 * (G.bar_id - 1) + k], (unsigned long long)pc.acc[k]);   // 32 slots: one set per group
}

template <bool UNIT_TEMP, int MODE>
__global__ void __launch_bounds__(2 * GT, 1) ac_duo_kernel(const __grid_constant__ ns_ac_params P, int32_t* slow_ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, grp = tid / GT, gt = tid % GT, warp = tid >> 5;
  double* tab = reinterpret_cast<double*>(smem_raw);
  DCta* cta = reinterpret_cast<DCta*>(smem_raw + NS_EXP_N * 8);
  unsigned char* gb = smem_raw + NS_EXP_N * 8 + D_CTA_BYTES + grp * D_GROUP_BYTES;
  DGroup G;
  G.tab = tab;
  G.hist = reinterpret_cast<uint32_t*>(gb);
  G.clist = reinterpret_cast<DCand*>(gb);
  G.band = reinterpret_cast<DBand*>(gb + D_NB * 4);
  G.ulist = reinterpret_cast<DUnd*>(gb + D_NB * 4 + D_BAND_CAP * 16);
  G.sc = reinterpret_cast<DScal*>(gb + D_NB * 4 + D_BAND_CAP * 16 + D_U_CAP * 8);
  G.words = reinterpret_cast<float*>(smem_raw + D_FIXED);
  G.bar_id = 1 + grp;
  for (int i = tid; i < NS_EXP_N; i += 2 * GT) tab[i] = c_exp_tab[i];
  if (tid == 0) cta->next_it = 0;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(d_saddr(&cta->tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = cta->tmem_base;
  // this thread's TMEM slots: the lanes of its warp's quadrant, one half of the 512 columns
  G.tmem = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 1) * 256);
  if (grp == 0) duo_group<UNIT_TEMP, MODE, STORE_SMEM>(P, slow_ws, G, cta, gt);
  else duo_group<UNIT_TEMP, MODE, STORE_TMEM>(P, slow_ws, G, cta, gt);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase));
}
