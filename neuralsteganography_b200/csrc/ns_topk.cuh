// ns_topk.cuh -- arithmetic-coder step, rank form of the cutoff (code_base/arithmetic.py:75 with top-k binding),
// WITHOUT a resident row (sm_100a).  Included by ns_coder.cu inside namespace nst, after the shared definitions.
//
// When at least topk tokens have p >= 1/range the kept set is the topk largest logits (arithmetic.py:142) and
// nothing below them matters: no fp64 work on the row at all, and the row is read from HBM exactly ONCE with a
// dozen instructions per 16-byte chunk.  One 512-thread CTA per row, 37 KB of shared memory, four CTAs per SM,
// so the chain of one row hides under the sweeps of the others.
//   sample   two chunks per thread (4096 logits, strided over the row): every warp takes the j-th largest of its 256
//            sample keys (j rounds of warp maximum + strike-out), the mean over the warps is a key bound k_c that about
//            2.5 topk + 64 keys of the row exceed (order statistics: relative spread 1/sqrt(16 j)).  Three threads of
//            three warps fetch what decides whether the row is coded (phase, slot, interval, message cursor) meanwhile
//   sweep    (HBM, once) per chunk: its maximum m -> row maximum M, an UPPER bound 4 sum_chunks exp((m - M)/temp)
//            of the softmax normaliser (one ex2 per chunk), and the chunk index to a hit list when m >= k_c
//   gather   the hit chunks again (L2): the keys >= k_c, the candidates; at least topk of them means the topk
//            largest keys of the row are the topk largest candidates (ties included: equal keys are candidates too).
//            The last three warps do not gather: one thread each derives the row constants (two of them) and fetches
//            the message window / the observed token
//   order    count histogram of the candidates over [k_c, M] -> grouped by bucket -> exact order inside each bucket
//            (key, lower id first) -> rank of every candidate; the row is certainly in rank form when the key of rank
//            topk-1 has p >= 1/range even against the upper bound of the normaliser (plus the guard band)
//   chain    exp64 of the topk kept tokens, bin widths, prefix sums, overfill, search (encode) / position of the
//            observed token (decode), interval update -- the arithmetic of fast_rank_row, same reduction order,
//            same integers
// Rows this kernel does not carry (not certainly in rank form, too few / too many candidates, degenerate ties,
// finish_sent tail) are queued in `mid` = {count, done, rows...}; ac_fast_kernel then runs on that list.

#ifndef NST_KT
#define NST_KT 512
#endif
constexpr int KT = NST_KT;                   // threads per CTA (512, or 256 with two ranks per thread in the chain)
constexpr int KW = KT / 32;
#ifndef NST_NB
#define NST_NB 2048
#endif
#ifndef NST_TOPK_CAP
#define NST_TOPK_CAP 512
#endif
constexpr int K_NB = NST_NB;                 // histogram buckets
constexpr int K_BPT = K_NB / KT;
#ifndef NST_CAP
#define NST_CAP 1792
#endif
constexpr int K_CAP = NST_CAP;               // candidates (keys >= k_c)
constexpr int K_HCAP = 2048;                 // chunks holding a candidate
constexpr int K_TOPK_CAP = NST_TOPK_CAP;              // one thread per kept token; 2.5 topk + 64 stays well below K_CAP
constexpr int K_MIN_VOCAB = 8192;            // the sample needs 1024 distinct interior chunks
#ifndef NST_U
#define NST_U 4
#endif
#ifndef NST_HINTS
#define NST_HINTS 1
#endif
#ifndef NST_MIN_CTAS
#define NST_MIN_CTAS 4
#endif
constexpr int K_U = NST_U;                   // chunk loads in flight per thread
constexpr int K_MIN_CTAS = NST_MIN_CTAS;     // CTAs per SM the register budget is cut for
constexpr int K_JMAX = 8;                    // rounds of the sample selection (the host sends smaller V / larger topk to ns_fast.cuh)
constexpr int K_TIE_CAP = 256;               // keys in the boundary bucket (the order inside a bucket is quadratic)
constexpr float K_MAGIC = 2097152.0f;        // 2^21: a float counts quarters there
constexpr float K_BAND_EPS = 0.0009765625f;  // the guard band of the throughput kernels (F_BAND_EPS)
static_assert(K_TOPK_CAP <= 2 * KT && K_TOPK_CAP * 12 <= K_CAP * 8, "chain arrays live in the candidate / list areas");
static_assert(K_HCAP * 4 <= K_CAP * 8, "the hit list lives in the list area");
static_assert(K_U >= 1 && K_U <= 8, "hit entries carry K_U mask bits under the chunk index");
static_assert(KT >= 256 && (KT & (KT - 1)) == 0, "chunk ownership by mask; three warps of helpers beside the gathering ones");
static_assert(5 * K_TOPK_CAP / 2 + 64 + 350 <= K_CAP, "room for the spread of the candidate count");

// shapes the sweep kernel takes: the sample selection needs its rank within K_JMAX rounds
inline bool k_shape_ok(int V, int K) {
  return V >= K_MIN_VOCAB && K >= 2 && K <= K_TOPK_CAP && (long long)((5 * K) / 2 + 64) * 256 + V / 2 < (long long)(K_JMAX + 1) * V;   // (256 sample keys per warp)
}
struct KCand { float key; int id; };
struct KScal {
  u64 red[KW > 16 ? KW : 16];                  // one per warp (scan) / per group of 32 ranks (chain)
  uint32_t wmax[KW], smax[KW], smin[KW];
  float wsum[KW];
  int sel_bin, res_idx, res_found, nhit, ncand;
  uint32_t sel_prefix, sel_cnt;
  float kth_key;
  // the stream's scalars (thread 0) and the row's constants (one thread, while the others gather)
  u64 lo, R, window;
  double dm;
  float M, kappa_r, clamp_key, scale2, boff2;
  int go, slot, cursor, mlen, tok, bad, bad2;
};

__device__ __forceinline__ float k_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// byte offset of the bucket of key k: (hi - k) * scale + 2^21 rounds to quarters; mantissa field with the two low bits
// masked = 4 * bucket (ns_codecs_stream.cuh, stream_off).  Monotone non-increasing in k, clamped to the histogram.
__device__ __forceinline__ uint32_t k_off(float k, float scale, float off) {
  return __float_as_uint(fmaxf(fminf(fmaf(-k, scale, off), K_MAGIC + (float)(K_NB - 1) + 0.75f), K_MAGIC)) & ((uint32_t)(K_NB - 1) << 2);
}
__device__ __forceinline__ void k_defer(const ns_ac_params& P, int32_t* mid, int row) {
  const int s = atomicAdd(&mid[0], 1);
  mid[2 + s] = row;
  if (P.status) atomicOr(&P.status[row], NS_ST_RANK_DEFER);   // informational
}

template <bool UNIT_TEMP, int MODE>
__global__ void __launch_bounds__(KT, K_MIN_CTAS) ac_topk_stream_kernel(const __grid_constant__ ns_ac_params P, int32_t* mid) {
  __shared__ __align__(16) uint32_t hist[K_NB];              // count histograms; then count | prefix << 16 per bucket
  __shared__ __align__(16) KCand list[K_CAP];                // hit chunks; candidates grouped by bucket; prefix sums of the bin widths
  __shared__ __align__(16) KCand cand[K_CAP];                // candidates in arrival order; later es[topk] (fp64) and sid[topk]
  __shared__ KScal sc;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, V = P.V, K = P.topk;
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  const int W4 = (mis + V + 3) >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g - mis);          // 16-byte aligned view
  const int mk0 = (P.mask_id[0] >= 0 && P.mask_id[0] < V) ? P.mask_id[0] : -8;
  const int mk1 = (P.mask_id[1] >= 0 && P.mask_id[1] < V) ? P.mask_id[1] : -8;
  const int mc0 = (mk0 + mis) >> 2, mc1 = (mk1 + mis) >> 2;
#if NST_HINTS
  // L2 eviction hints: the sweep's lines may leave first, the sample's (read again by the sweep) stay.  The policy is made
  // at run time from a fraction the compiler cannot fold -- a compile-time policy is rebuilt in every loop iteration.
  u64 pol_last, pol_first;
  {
    const float one = __uint_as_float(0x3f800000u | ((uint32_t)P.B >> 31));      // 1.0f (B >= 0)
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(pol_last) : "f"(one));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, %1;" : "=l"(pol_first) : "f"(one));
  }
  auto ldg4 = [&](int c, bool last) -> float4 {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(g4 + c), "l"(last ? pol_first : pol_last));
    return v;
  };
#else
  // (no L2 eviction hints: a policy operand costs the sweep loop a descriptor set-up per load, and the few chunks
  // that are read twice -- the sample, the hit chunks -- come back from L2 within microseconds anyway)
  auto ldg4 = [&](int c, bool) -> float4 {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(g4 + c));
    return v;
  };
#endif
  // slot counter in shared memory, one plain atomic per calling thread (the compiler's warp-aggregated form of
  // atomicAdd costs twenty instructions on a path nearly every warp takes for two or three of its lanes)
  auto take_slot = [&](int* counter) -> int {
    int old;
    asm volatile("atom.shared.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(old) : "r"((uint32_t)__cvta_generic_to_shared(counter)) : "memory");   // (inc: ptxas turns add-1 into the aggregated form)
    return old;
  };
  // forbidden tokens (arithmetic.py:124-125): probability exactly 0 -- the logits are read-only, so on the fly
  auto fold = [&](float4 v, int c) -> float4 {
    if (c == mc0 || c == mc1) {
      const int b = 4 * c - mis;
      float* f = reinterpret_cast<float*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) if (b + j == mk0 || b + j == mk1) f[j] = -INFINITY;
    }
    return v;
  };
  // an edge chunk element-wise (-inf outside the row)
  auto edge = [&](int c) -> float4 {
    const int b0 = 4 * c - mis;
    float4 v;
    v.x = (b0 >= 0 && b0 < V) ? g[b0] : -INFINITY;
    v.y = (b0 + 1 >= 0 && b0 + 1 < V) ? g[b0 + 1] : -INFINITY;
    v.z = (b0 + 2 >= 0 && b0 + 2 < V) ? g[b0 + 2] : -INFINITY;
    v.w = (b0 + 3 >= 0 && b0 + 3 < V) ? g[b0 + 3] : -INFINITY;
    return v;
  };
  // exclusive scan of the histogram (bucket 0 = largest keys) and the bucket holding position `pos`; the buckets
  // become count | prefix << 16 (pack) or zero
  auto scan_find = [&](int pos, bool pack) {
    uint32_t hl[K_BPT], tsum = 0;
#pragma unroll
    for (int b = 0; b < K_BPT; ++b) { hl[b] = hist[tid * K_BPT + b]; tsum += hl[b]; }
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sc.red[warp] = (u64)inc;
    if (tid == 0) { sc.sel_bin = -1; sc.sel_prefix = 0; sc.sel_cnt = 0; sc.res_idx = K; sc.res_found = 0; }
    __syncthreads();
    uint32_t excl = inc - tsum;
#pragma unroll
    for (int w = 0; w < KW; ++w) if (w < warp) excl += (uint32_t)sc.red[w];
#pragma unroll
    for (int b = 0; b < K_BPT; ++b) {
      const uint32_t c = hl[b];
      if (c != 0 && excl <= (uint32_t)pos && (uint32_t)pos < excl + c) { sc.sel_bin = tid * K_BPT + b; sc.sel_prefix = excl; sc.sel_cnt = c; }
      hist[tid * K_BPT + b] = pack ? (c | (excl << 16)) : 0u;
      excl += c;
    }
    __syncthreads();
  };
  for (int i = tid; i < K_NB; i += KT) hist[i] = 0;
  // ---- sample: two chunks per thread, strided over the row -> reference of the bound, key bound k_c
  float k_c, ref;
  {
    const int NI = W4 - 2;                                   // interior chunks 1 .. W4-2 (>= 2 KT: K_MIN_VOCAB)
    const int cs0 = 1 + (int)(((long long)tid * NI) / (2 * KT)), cs1 = 1 + (int)(((long long)(tid + KT) * NI) / (2 * KT));
    float4 a = ldg4(cs0, false), b = ldg4(cs1, false);
    // ---- the stream's scalars: three threads of three warps (loads issued together, behind the sample loads), handed
    // on through shared memory at the first barrier.  Only what decides whether the row is coded and where its state
    // lies is fetched here; what depends on it (message window, observed token, 1/range) is fetched while the
    // candidates are gathered, by warps that do not gather
    if (tid == 0) {                                          // may the row be coded at all; slot, observed token
      const int ph0 = P.phase ? (int)P.phase[row] : NS_PHASE_CODING;
      const int slot = P.ntok ? P.ntok[row] : 0;
      int ntot = 0x7fffffff;
      if (MODE == MODE_DEC && P.ntok_total) ntot = P.ntok_total[row];
      int go = 1;
      if (ph0 == NS_PHASE_DONE) go = 0;
      const int phase = MODE == MODE_ENC ? ph0 : NS_PHASE_CODING;
      if (go && MODE == MODE_ENC && P.ntok && slot >= P.token_cap) {
        if (P.phase) P.phase[row] = NS_PHASE_DONE;
        if (P.status) atomicOr(&P.status[row], NS_ST_TOKEN_OVERFLOW);
        go = 0;
      }
      if (go && MODE == MODE_DEC && P.ntok_total && slot >= ntot) {
        if (P.phase) P.phase[row] = NS_PHASE_DONE;
        go = 0;
      }
      if (go && MODE == MODE_ENC && phase == NS_PHASE_TAIL) {  // finish_sent tail (:135-137): the row-resident kernel emits rank 0
        k_defer(P, mid, row);
        go = 0;
      }
      sc.go = go; sc.slot = slot;
      sc.nhit = 0; sc.ncand = 0; sc.kth_key = -INFINITY;
    } else if (tid == 32) {                                  // the interval
      const u64 lo = P.lo[row], R = P.hi[row] - lo;          // arithmetic.py:140
      sc.lo = lo; sc.R = R;
    } else if (tid == 64 && MODE == MODE_ENC) {              // where the message stands
      sc.cursor = P.cursor[row]; sc.mlen = P.msg_len[row];
    }
    a = fold(a, cs0); b = fold(b, cs1);
    // j-th largest of the warp's 256 sample keys (ties struck out together), j ~ (2.5 K + 64) 256 / V: about
    // 2.5 K + 64 keys of the row exceed it; the mean over the warps has a sixteenth of one warp's variance
    uint32_t ks[8] = {ns_f32_orderable(a.x), ns_f32_orderable(a.y), ns_f32_orderable(a.z), ns_f32_orderable(a.w),
                      ns_f32_orderable(b.x), ns_f32_orderable(b.y), ns_f32_orderable(b.z), ns_f32_orderable(b.w)};
    int jr = (((5 * K) / 2 + 64) * (8 * 32) + V / 2) / V;
    jr = jr < 1 ? 1 : (jr > K_JMAX ? K_JMAX : jr);
    uint32_t top = 0, cur = 0;
    for (int r = 0; r < jr; ++r) {
      uint32_t m = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) m = max(m, ks[i]);
      cur = __reduce_max_sync(0xffffffffu, m);
      if (r == 0) top = cur;
#pragma unroll
      for (int i = 0; i < 8; ++i) ks[i] = ks[i] == cur ? 0u : ks[i];
    }
    if (lane == 0) { sc.smax[warp] = top; sc.smin[warp] = cur; }
    __syncthreads();                                         // (also: the histogram is clear, the stream's scalars are there)
    if (!sc.go) return;
    uint32_t ga = 0;
    float ksum = 0.f;
#pragma unroll
    for (int w = 0; w < KW; ++w) { ga = max(ga, sc.smax[w]); ksum += key_of_pack((u64)sc.smin[w] << 32); }
    ref = key_of_pack((u64)ga << 32);
    k_c = ksum * (1.0f / (float)KW);
    if (!(ref > -3.0e38f) || !(ref < 3.0e38f) || !(k_c > -3.0e38f) || !(k_c < 3.0e38f)) {
      if (tid == 0) k_defer(P, mid, row);
      return;
    }
  }
  // ---- the sweep (HBM, once): chunk maxima -> row maximum, upper bound of the normaliser, hit list
  const float c2 = (float)(1.4426950408889634 / P.temp);     // log2(e)/temp
  int* hits = reinterpret_cast<int*>(list);                  // [K_HCAP]
  {
    float bk = -INFINITY, ts0 = 0.f, ts1 = 0.f;
    const float nrc = -ref * c2;
    auto one = [&](const float4 v, float& ts) -> bool {
      const float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      bk = fmaxf(bk, m);
      ts += k_ex2(fmaf(m, c2, nrc));
      return m >= k_c;
    };
    // hit list entry: first chunk << K_U | mask of the thread's chunks c + u KT holding a candidate
    auto push = [&](int c, uint32_t hm) {
      const int p = take_slot(&sc.nhit);
      if (p < K_HCAP) hits[p] = (int)(((uint32_t)c << K_U) | hm);
    };
    // interior chunk c belongs to thread (c - 1) mod KT; the (at most two) chunks with a forbidden id are folded by
    // their owner in the iteration that holds them -- one range test per K_U chunks
    const int own0 = (mc0 >= 1 && mc0 < W4 - 1 && ((mc0 - 1) & (KT - 1)) == tid) ? mc0 : -(1 << 30);
    const int own1 = (mc1 >= 1 && mc1 < W4 - 1 && ((mc1 - 1) & (KT - 1)) == tid) ? mc1 : -(1 << 30);
    int c = 1 + tid;
    for (; c + (K_U - 1) * KT < W4 - 1; c += K_U * KT) {
      float4 v[K_U];
#pragma unroll
      for (int u = 0; u < K_U; ++u) v[u] = ldg4(c + u * KT, true);
      if ((unsigned)(own0 - c) < (unsigned)(K_U * KT) || (unsigned)(own1 - c) < (unsigned)(K_U * KT)) {
#pragma unroll
        for (int u = 0; u < K_U; ++u) v[u] = fold(v[u], c + u * KT);
      }
      uint32_t hm = 0;
#pragma unroll
      for (int u = 0; u < K_U; ++u) hm |= one(v[u], (u & 1) ? ts1 : ts0) ? (1u << u) : 0u;
      if (hm) push(c, hm);
    }
    for (; c < W4 - 1; c += KT) if (one(fold(ldg4(c, true), c), ts0)) push(c, 1u);
    if (tid == 0 || tid == 32) { const int ce = tid ? W4 - 1 : 0; if (one(fold(edge(ce), ce), ts1)) push(ce, 1u); }
    const uint32_t wk = __reduce_max_sync(0xffffffffu, ns_f32_orderable(bk + 0.0f));
    float wts = ts0 + ts1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wts += __shfl_xor_sync(0xffffffffu, wts, o);
    if (lane == 0) { sc.wmax[warp] = wk; sc.wsum[warp] = wts; }
  }
  __syncthreads();                                           // the hit list and the warps' maxima / sums are complete
  // ---- row constants and the stream's dependent scalars: three threads of the last three warps (the formulas of
  // fast_row, with the bound in place of the estimate) ...
  constexpr int K_GT = KT - 96;                              // threads that gather meanwhile
  if (tid == KT - 32) {                                      // bound of the normaliser -> key above which p >= 1/R for certain
    uint32_t mk = 0;
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < KW; ++w) { mk = max(mk, sc.wmax[w]); tot += sc.wsum[w]; }
    const float M = key_of_pack((u64)mk << 32);
    const float ssum = 4.0f * tot * k_ex2((ref - M) * c2);   // >= sum_j exp((x_j - M)/temp): four keys per chunk, each <= its maximum
    const float tf = (float)P.temp;
    const u64 R = sc.R;
    const double thr = __ddiv_rn(1.0, (double)R);            // :141
    const float key_th = fmaf(tf * 0.6931471805599453f, __log2f((float)(thr * (double)ssum)), M);   // p >= 1/R <= key >= M + temp ln(bound / R)
    const float clamp_key = (float)((double)M - 700.0 * P.temp);
    sc.kappa_r = key_th + tf * K_BAND_EPS + 0.02f * tf;      // guard for the fp32 arithmetic of the bound
    sc.bad = (!(ssum > 0.0f) || !(ssum < 3.0e38f) || !(R >= 2) || !(key_th > clamp_key) || !(M > -3.0e38f) || !(M < 3.0e38f)) ? 1 : 0;
  } else if (tid == KT - 64) {                               // maximum -> bucket function of the candidates, exponent offset
    uint32_t mk = 0;
#pragma unroll
    for (int w = 0; w < KW; ++w) mk = max(mk, sc.wmax[w]);
    const float M = key_of_pack((u64)mk << 32);
    const double Md = (double)M;
    const float scale2 = (float)K_NB / (M - k_c);
    const float boff2 = M * scale2 + K_MAGIC;
    sc.M = M;
    sc.dm = UNIT_TEMP ? Md : __ddiv_rn(Md, P.temp);
    sc.clamp_key = (float)(Md - 700.0 * P.temp);
    sc.scale2 = scale2; sc.boff2 = boff2;
    sc.bad2 = (!(M > k_c) || !(scale2 > 0.0f) || !(scale2 < 3.0e38f) || !(fabsf(boff2) < 3.0e38f)) ? 1 : 0;
  } else if (tid == KT - 96) {                               // the next `precision` message bits (:168-171) / the observed token
    if (MODE == MODE_ENC) sc.window = ns_read_bits(P.msg + (size_t)row * P.msg_stride, sc.cursor, sc.mlen, P.precision);
    else sc.tok = P.token_in[(size_t)row * P.token_stride + sc.slot];
  }
  // ---- ... while the other warps gather (L2) the candidates = keys >= k_c of the hit chunks
  if (tid < K_GT) {
    const int nhit = min(sc.nhit, K_HCAP);
    for (int h = tid; h < nhit; h += K_GT) {
      const uint32_t ent = (uint32_t)hits[h];
      for (uint32_t hm = ent & ((1u << K_U) - 1u); hm != 0; hm &= hm - 1u) {
        const int c = (int)(ent >> K_U) + (__ffs((int)hm) - 1) * KT, b0 = 4 * c - mis;
        const float4 v = (c >= 1 && c < W4 - 1) ? ldg4(c, true) : edge(c);
        const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (xs[j] >= k_c && (unsigned)(b0 + j) < (unsigned)V && b0 + j != mk0 && b0 + j != mk1) {
            const int p = take_slot(&sc.ncand);
            if (p < K_CAP) { KCand e; e.key = xs[j] + 0.0f; e.id = b0 + j; cand[p] = e; }
          }
      }
    }
  }
  __syncthreads();
  const int nc = sc.ncand;
  // not carried: constants out of range, hit list full, the sample bound missed (rare) or ties past the list
  if (sc.bad || sc.bad2 || sc.nhit > K_HCAP || nc < K || nc > K_CAP) {
    if (tid == 0) k_defer(P, mid, row);
    return;
  }
  // ---- order: count histogram of the candidates over [k_c, M], grouped by bucket, exact order inside each bucket
  const float scale2 = sc.scale2, boff2 = sc.boff2;
  {
    const uint32_t hb = (uint32_t)__cvta_generic_to_shared(hist);
    for (int i = tid; i < nc; i += KT)
      asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hb + k_off(cand[i].key, scale2, boff2)), "r"(1u) : "memory");
  }
  __syncthreads();
  scan_find(K - 1, true);                                    // bucket of the last kept position
  const int total = (int)(sc.sel_prefix + sc.sel_cnt);       // candidates in the buckets up to that one: at least K
  if (sc.sel_bin < 0 || sc.sel_cnt > (uint32_t)K_TIE_CAP) {
    if (tid == 0) k_defer(P, mid, row);
    return;
  }
  for (int i = tid; i < nc; i += KT) {
    const KCand cnd = cand[i];
    const uint32_t off = k_off(cnd.key, scale2, boff2);
    const uint32_t old = atomicSub(&hist[off >> 2], 1u);     // low 16 bits: slots still free in the bucket
    list[(old >> 16) + (old & 0xffffu) - 1u] = cnd;
  }
  __syncthreads();                                           // (the buckets now hold prefix << 16: where each bucket starts)
  double* es = reinterpret_cast<double*>(cand);              // [K_TOPK_CAP]
  int* sid = reinterpret_cast<int*>(cand + K_TOPK_CAP);      // [K_TOPK_CAP]
  {
    const float clamp_key = sc.clamp_key;
    const double dm = sc.dm;
    for (int p = tid; p < total; p += KT) {
      const KCand me = list[p];
      const uint32_t b = k_off(me.key, scale2, boff2) >> 2;
      const int ex = (int)(hist[b] >> 16), end = b + 1 < (uint32_t)K_NB ? (int)(hist[b + 1] >> 16) : nc;
      int r = ex;
      for (int o = ex; o < end; ++o) {
        const KCand ot = list[o];
        r += (ot.key > me.key || (ot.key == me.key && ot.id < me.id)) ? 1 : 0;   // coder order: key, then lower id
      }
      if (r < K) {
        double x = (double)fmaxf(me.key, clamp_key);         // (double(x)/temp) - (double(max)/temp), :128-130
        if (!UNIT_TEMP) x = __ddiv_rn(x, P.temp);
        es[r] = ns_exp64_core(x - dm, g_exp_tab);           // (table in global memory: every lane its own index)
        sid[r] = me.id;
        if (r == K - 1) sc.kth_key = me.key;
      }
    }
  }
  __syncthreads();
  if (!(sc.kth_key >= sc.kappa_r)) {                         // top-k does not certainly bind (:75): the threshold form decides
    if (tid == 0) k_defer(P, mid, row);
    return;
  }
  // ---- bin widths, prefix sums, overfill, selection: thread t holds the tokens of rank t + s KT (s < RPT); only the
  // warps that hold a rank go on (named barrier over those warps).  Groups of 32 consecutive ranks are reduced by one
  // warp each and combined in rank order: the reduction order of fast_rank_row, whatever KT is.
  constexpr int RPT = (K_TOPK_CAP + KT - 1) / KT;            // ranks per thread
  static_assert(RPT * KW <= (KW > 16 ? KW : 16), "one reduction slot per group of 32 ranks");
  const int NG = (K + 31) >> 5;                              // groups that hold a rank
  const int nwt = NG * 32 < KT ? NG * 32 : KT;               // threads of the warps holding a rank
  if (tid >= nwt) return;
  auto tail_sync = [&]() { asm volatile("bar.sync 1, %0;" :: "r"(nwt) : "memory"); };
  const u64 lo = sc.lo, R = sc.R;
  double ev[RPT];
#pragma unroll
  for (int s = 0; s < RPT; ++s) {
    const int r = tid + s * KT;
    ev[s] = r < K ? es[r] : 0.0;
    double S = ev[s];                                        // sum of the kept e, fixed order (:146)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) S = S + __shfl_xor_sync(0xffffffffu, S, o);
    if (lane == 0) sc.red[warp + s * KW] = (u64)__double_as_longlong(S);
  }
  tail_sync();
  double S = __longlong_as_double((long long)sc.red[0]);
  for (int g = 1; g < NG; ++g) S = S + __longlong_as_double((long long)sc.red[g]);   // groups without a rank would add 0.0: same bits
  const double C = __ddiv_rn((double)R, S);
  u64 cum[RPT];
#pragma unroll
  for (int s = 0; s < RPT; ++s) {
    const int r = tid + s * KT;
    cum[s] = r < K ? (u64)__double2ll_rn(ev[s] * C) : 0ull;  // :146-149; inclusive prefix sums over the ranks (:150)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, cum[s], o); if (lane >= o) cum[s] += t; }
  }
  tail_sync();                                               // red of the sum is consumed
#pragma unroll
  for (int s = 0; s < RPT; ++s) if (lane == 31) sc.red[warp + s * KW] = cum[s];
  tail_sync();
  u64 Q = 0;
  {
    u64 wo[RPT];
#pragma unroll
    for (int s = 0; s < RPT; ++s) wo[s] = 0;
    for (int g = 0; g < NG; ++g) {
      const u64 x = sc.red[g];
#pragma unroll
      for (int s = 0; s < RPT; ++s) if (g < warp + s * KW) wo[s] += x;
      Q += x;
    }
#pragma unroll
    for (int s = 0; s < RPT; ++s) cum[s] += wo[s];
  }
  u64* cums = reinterpret_cast<u64*>(list);                  // [K_TOPK_CAP]; the gathered list is no longer needed
#pragma unroll
  for (int s = 0; s < RPT; ++s) if (tid + s * KT < K) cums[tid + s * KT] = cum[s];
  tail_sync();
  // overfill (:153-158): drop the ranks from the first prefix sum above the range on
  int kk = K;
  u64 slack;
  if (Q > R) {
#pragma unroll
    for (int s = 0; s < RPT; ++s) {
      const int r = tid + s * KT;
      if (r < K && cum[s] > R && (r == 0 || cums[r - 1] <= R)) sc.res_idx = r;
    }
    tail_sync();
    kk = sc.res_idx;
    slack = R - (kk > 0 ? cums[kk - 1] : 0ull);
    tail_sync();
    if (tid == 0) sc.res_idx = K;
    tail_sync();
  } else {
    slack = R - Q;
  }
  // bin of rank r: [cums[r-1] + slack, cums[r] + slack), rank 0 starts at 0 and absorbs the slack (:158)
  u64 my_lo[RPT], my_hi[RPT];
#pragma unroll
  for (int s = 0; s < RPT; ++s) {
    const int r = tid + s * KT;
    my_lo[s] = (r > 0 && r < K) ? cums[r - 1] + slack : 0ull;
    my_hi[s] = cum[s] + slack;
  }
  if (MODE == MODE_ENC) {
    const u64 m_rel = sc.window - lo;                        // next `precision` message bits (:168-171)
#pragma unroll
    for (int s = 0; s < RPT; ++s)
      if (tid + s * KT < kk && my_lo[s] <= m_rel && m_rel < my_hi[s]) sc.res_idx = tid + s * KT;   // :172 (empty bins never match)
    tail_sync();
    const int r = sc.res_idx;
    const int rr = r < kk ? r : 0;
#pragma unroll
    for (int s = 0; s < RPT; ++s)
      if (tid + s * KT == rr) {
        if (r >= kk && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW);   // cannot happen: the bins tile the range
        finish_encode(P, row, sc.slot, sid[rr], lo + my_lo[s], lo + my_hi[s], (u64)K, Q, sc.cursor, sc.mlen);   // :175-176
      }
  } else {
    const int tok = sc.tok;
    const bool tok_ok = tok >= 0 && tok < V;
#pragma unroll
    for (int s = 0; s < RPT; ++s)
      if (tid + s * KT < kk && tok_ok && sid[tid + s * KT] == tok) { sc.res_idx = tid + s * KT; sc.res_found = 1; }
    tail_sync();
    const bool in_range = sc.res_found != 0;
    const int r = in_range ? sc.res_idx : 0;                 // :342 / :347-348: unknown tokens are coded as rank 0
#pragma unroll
    for (int s = 0; s < RPT; ++s)
      if (tid + s * KT == r) finish_decode(P, row, sc.slot, in_range || !tok_ok, lo + my_lo[s], lo + my_hi[s], (u64)K, Q);
  }
}
