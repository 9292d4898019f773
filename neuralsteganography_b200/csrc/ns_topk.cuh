// ns_topk.cuh -- arithmetic-coder step, rank form of the cutoff (code_base/arithmetic.py:75 with top-k binding),
// WITHOUT a resident row (sm_100a).  Included by ns_coder.cu inside namespace nst, after the shared definitions.
//
// When more than topk tokens have p >= 1/range the kept set is the topk largest logits (arithmetic.py:142) and
// nothing below them matters: no fp64 work on the row at all.  The row-resident kernel (ns_fast.cuh,
// fast_rank_row) spends its time in a chain of ~25 barrier-separated steps on one row per SM (issue slots 36 %
// busy).  Here the row is swept from global memory like the rank codec does (ns_codecs_stream.cuh): one
// 512-thread CTA per row, 25 KB of shared memory, two CTAs per SM, so the chain of one row hides under the
// sweeps of another.
//   sample   one chunk per thread: bucket range of the count histogram (any monotone bucket function gives
//            the same kept set: the order inside a bucket is resolved exactly) and the reference of the estimate
//   sweep 1  (HBM) row maximum, fp32 estimate of sum exp((x - M)/temp), count histogram of the keys
//   scan     bucket holding position topk of the coder's order; the row is certainly in rank form when every
//            key of the buckets up to that one lies above the estimated cutoff plus the 2 % guard
//   sweep 2  (L2) the keys of those buckets, grouped by bucket at their prefix (count | prefix per bucket)
//   chain    order inside each bucket (key, lower id first), exp64 of the topk kept tokens, bin widths,
//            prefix sums, overfill, search (encode) / position of the observed token (decode), interval update
//            -- the arithmetic of fast_rank_row, same reduction order, same integers
// Rows this kernel does not carry (not certainly in rank form, estimate outside its guard, lists too small,
// finish_sent tail) are queued in `mid` = {count, done, rows...}; ac_fast_kernel then runs on that list.

constexpr int KT = 512;                      // threads per CTA
constexpr int KW = KT / 32;
constexpr int K_NB = 2048;                   // histogram buckets
constexpr int K_BPT = K_NB / KT;
constexpr int K_CAP = 1024;                  // gathered keys (the topk kept ones + the rest of the boundary bucket)
constexpr int K_TOPK_CAP = 512;              // one thread per kept token
constexpr int K_MIN_VOCAB = 4096;
constexpr float K_MAGIC = 2097152.0f;        // 2^21: a float counts quarters there
constexpr float K_BAND_EPS = 0.0009765625f;  // the guard band of the throughput kernels (F_BAND_EPS)
template <int N> struct KDepth { static constexpr int value = N; };   // chunk loads in flight per thread of a sweep
static_assert(K_TOPK_CAP <= KT && K_TOPK_CAP * 12 <= K_NB * 4 && K_TOPK_CAP * 8 <= K_CAP * 16, "chain arrays live in the histogram / list areas");

struct KEntry { float key; int id; uint32_t ex; uint32_t pad; };   // ex = first list position of the entry's bucket
struct KScal {
  u64 red[KW];
  uint32_t wmax[KW], smax[KW], smin[KW];
  float wsum[KW];
  int sel_bin, res_idx, res_found, pad;
  uint32_t sel_prefix, sel_cnt;
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
__global__ void __launch_bounds__(KT, 2) ac_topk_stream_kernel(const __grid_constant__ ns_ac_params P, int32_t* mid) {
  __shared__ __align__(16) uint32_t hist[K_NB];              // count histogram; later es[topk] (fp64) and sid[topk]
  __shared__ __align__(16) KEntry list[K_CAP];               // gathered keys; later the prefix sums of the bin widths
  __shared__ KScal sc;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, V = P.V, K = P.topk;
  // ---- the stream's scalars
  int phase = P.phase ? (int)P.phase[row] : NS_PHASE_CODING;
  if (phase == NS_PHASE_DONE) return;
  if (MODE != MODE_ENC) phase = NS_PHASE_CODING;
  const int slot = P.ntok ? P.ntok[row] : 0;
  if (MODE == MODE_ENC && P.ntok && slot >= P.token_cap) {
    if (tid == 0) { if (P.phase) P.phase[row] = NS_PHASE_DONE; if (P.status) atomicOr(&P.status[row], NS_ST_TOKEN_OVERFLOW); }
    return;
  }
  if (MODE == MODE_DEC && P.ntok_total && slot >= P.ntok_total[row]) {
    if (tid == 0 && P.phase) P.phase[row] = NS_PHASE_DONE;
    return;
  }
  if (MODE == MODE_ENC && phase == NS_PHASE_TAIL) {          // finish_sent tail (:135-137): the row-resident kernel emits rank 0
    if (tid == 0) k_defer(P, mid, row);
    return;
  }
  const u64 lo = P.lo[row], R = P.hi[row] - lo;              // arithmetic.py:140
  int cursor = 0, mlen = 0, tok = -1;
  u64 window = 0;
  if (MODE == MODE_ENC) {
    cursor = P.cursor[row]; mlen = P.msg_len[row];
    window = ns_read_bits(P.msg + (size_t)row * P.msg_stride, cursor, mlen, P.precision);   // :168-171
  } else {
    tok = P.token_in[(size_t)row * P.token_stride + slot];
  }
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  const int W4 = (mis + V + 3) >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g - mis);          // 16-byte aligned view
  const int mk0 = (P.mask_id[0] >= 0 && P.mask_id[0] < V) ? P.mask_id[0] : -8;
  const int mk1 = (P.mask_id[1] >= 0 && P.mask_id[1] < V) ? P.mask_id[1] : -8;
  const int mc0 = (mk0 + mis) >> 2, mc1 = (mk1 + mis) >> 2;
  u64 pol_last, pol_first;                                   // L2 eviction policies for the row's lines
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  auto ldg4 = [&](int c, bool last) -> float4 {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(g4 + c), "l"(last ? pol_first : pol_last));
    return v;
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
  // interior chunks 1 .. W4-2 with U straight 128-bit loads in flight per thread, then the two edge chunks
  // element-wise by two threads (-inf outside the row).  body(v, first id, edge)
  auto sweep = [&](bool last, auto body, auto depth) {
    constexpr int U = decltype(depth)::value;
    int c = 1 + tid;
    for (; c + (U - 1) * KT < W4 - 1; c += U * KT) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ldg4(c + u * KT, last);
#pragma unroll
      for (int u = 0; u < U; ++u) body(fold(v[u], c + u * KT), 4 * (c + u * KT) - mis, false);
    }
    for (; c < W4 - 1; c += KT) body(fold(ldg4(c, last), c), 4 * c - mis, false);
    if (tid == 0 || tid == 32) {
      const int ce = tid ? W4 - 1 : 0, b0 = 4 * ce - mis;
      float4 v;
      v.x = (b0 >= 0 && b0 < V) ? g[b0] : -INFINITY;
      v.y = (b0 + 1 >= 0 && b0 + 1 < V) ? g[b0 + 1] : -INFINITY;
      v.z = (b0 + 2 >= 0 && b0 + 2 < V) ? g[b0 + 2] : -INFINITY;
      v.w = (b0 + 3 >= 0 && b0 + 3 < V) ? g[b0 + 3] : -INFINITY;
      body(fold(v, ce), b0, true);
    }
  };
  for (int i = tid; i < K_NB; i += KT) hist[i] = 0;
  // ---- sample: extent of one chunk per thread -> bucket range (widened: the sample misses the tails), reference
  float scale, boff, ref;
  {
    float sb = -INFINITY, sl = INFINITY;
    const int stride = (W4 - 2) / KT > 0 ? (W4 - 2) / KT : 1;
    const int cs = 1 + tid * stride;
    if (cs < W4 - 1) {
      const float4 v = fold(ldg4(cs, false), cs);
      sb = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      sl = fminf(fminf(v.x > -1e9f ? v.x : INFINITY, v.y > -1e9f ? v.y : INFINITY),
                 fminf(v.z > -1e9f ? v.z : INFINITY, v.w > -1e9f ? v.w : INFINITY));
    }
    const uint32_t a = __reduce_max_sync(0xffffffffu, ns_f32_orderable(sb));
    const uint32_t b = __reduce_min_sync(0xffffffffu, ns_f32_orderable(sl));
    if (lane == 0) { sc.smax[warp] = a; sc.smin[warp] = b; }
    __syncthreads();                                         // (also: the histogram is clear)
    uint32_t ga = 0, gb = 0xffffffffu;
#pragma unroll
    for (int w = 0; w < KW; ++w) { ga = max(ga, sc.smax[w]); gb = min(gb, sc.smin[w]); }
    float smax = key_of_pack((u64)ga << 32), smin = key_of_pack((u64)gb << 32);
    if (!(smax > -3.0e38f) || !(smax < 3.0e38f) || !(smin > -3.0e38f) || !(smin < 3.0e38f)) { smax = 1.0f; smin = -1.0f; }
    float span = smax - smin;
    if (!(span > 0.0f)) span = 1.0f;
    const float hi_p = smax + 0.25f * span, lo_p = smin - 0.25f * span;
    scale = (float)K_NB / (hi_p - lo_p);
    boff = hi_p * scale + K_MAGIC;
    if (!(scale > 0.0f) || !(scale < 3.0e38f) || !(fabsf(boff) < 3.0e38f)) { scale = 0.0f; boff = K_MAGIC; }
    ref = smax;
  }
  // ---- sweep 1 (HBM): maximum, estimate of the softmax normaliser against the fixed reference, count histogram
  const float c2 = (float)(1.4426950408889634 / P.temp);     // log2(e)/temp
  float bk = -INFINITY, ts0 = 0.f, ts1 = 0.f;
  {
    const float nrc = -ref * c2;
    const uint32_t hb = (uint32_t)__cvta_generic_to_shared(hist);
    sweep(false, [&](const float4 v, int id, bool edge) {
      bk = fmaxf(bk, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
      ts0 += k_ex2(fmaf(v.x, c2, nrc)); ts1 += k_ex2(fmaf(v.y, c2, nrc));
      ts0 += k_ex2(fmaf(v.z, c2, nrc)); ts1 += k_ex2(fmaf(v.w, c2, nrc));
      const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (edge && (unsigned)(id + j) >= (unsigned)V) continue;          // padding of the edge chunks is not a token
        asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hb + k_off(xs[j], scale, boff)), "r"(1u) : "memory");
      }
    }, KDepth<8>());
  }
  float M, ssum;
  {
    const uint32_t wk = __reduce_max_sync(0xffffffffu, ns_f32_orderable(bk + 0.0f));
    float wts = ts0 + ts1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wts += __shfl_xor_sync(0xffffffffu, wts, o);
    if (lane == 0) { sc.wmax[warp] = wk; sc.wsum[warp] = wts; }
    __syncthreads();                                         // (also: the histogram is complete)
    uint32_t mk = 0;
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < KW; ++w) { mk = max(mk, sc.wmax[w]); tot += sc.wsum[w]; }
    M = key_of_pack((u64)mk << 32);
    ssum = tot * k_ex2((ref - M) * c2);
  }
  // ---- row constants (every thread, same bits): the formulas of fast_row
  const double thr = __ddiv_rn(1.0, (double)R);              // :141
  const double Md = (double)M;
  const double dm = UNIT_TEMP ? Md : __ddiv_rn(Md, P.temp);
  const float tf = (float)P.temp;
  const float key_th = fmaf(tf * 0.6931471805599453f, __log2f((float)(thr * (double)ssum)), M);   // p >= 1/R <=> key >= M + temp ln(sum / R)
  const float kappa_lo = key_th - tf * K_BAND_EPS;
  const float kappa_r = key_th + tf * K_BAND_EPS + 0.02f * tf;   // top-k binds even should the estimate be 2 % off
  const float clamp_key = (float)(Md - 700.0 * P.temp);
  if (!(ssum > 0.0f) || !(ssum < 3.0e38f) || !(R >= 2) || !(kappa_lo > clamp_key) || !(M > -3.0e38f) || !(M < 3.0e38f) || !(scale > 0.0f)) {
    if (tid == 0) k_defer(P, mid, row);
    return;
  }
  // ---- scan: bucket of position K (0-based: the first token NOT kept); buckets become count | prefix << 16
  {
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
      if (c != 0 && excl <= (uint32_t)K && (uint32_t)K < excl + c) { sc.sel_bin = tid * K_BPT + b; sc.sel_prefix = excl; sc.sel_cnt = c; }
      hist[tid * K_BPT + b] = c | ((excl < 0xffffu ? excl : 0xffffu) << 16);
      excl += c;
    }
    __syncthreads();
  }
  const int tb = sc.sel_bin;
  const int total = (int)(sc.sel_prefix + sc.sel_cnt);       // keys in the buckets 0 .. tb: at least K + 1
  // every key of bucket b satisfies (hi_p - k) scale < b + 7/8; one more bucket absorbs the rounding of the offset
  const float hi_p = (boff - K_MAGIC) / scale;
  const float key_lb = hi_p - ((float)tb + 2.0f) / scale;
  const float key_gather = hi_p - ((float)tb + 3.0f) / scale;   // chunk test of the gather: below every key of bucket tb
  if (tb < 0 || tb >= K_NB - 1 || total > K_CAP || !(key_lb >= kappa_r)) {   // not certainly in rank form / too many ties
    if (tid == 0) k_defer(P, mid, row);
    return;
  }
  // ---- sweep 2 (L2): the keys of buckets 0 .. tb, grouped by bucket
  {
    const uint32_t tboff = (uint32_t)tb << 2;
    sweep(true, [&](const float4 v, int id, bool) {
      if (fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) >= key_gather) {
        const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = k_off(xs[j], scale, boff);
          if (off <= tboff && (unsigned)(id + j) < (unsigned)V) {
            const uint32_t old = atomicSub(&hist[off >> 2], 1u);        // low 16 bits: slots still free in the bucket
            const uint32_t ex = old >> 16, pos = ex + (old & 0xffffu) - 1u;
            if (pos < (uint32_t)K_CAP) { KEntry e; e.key = xs[j] + 0.0f; e.id = id + j; e.ex = ex; e.pad = 0u; list[pos] = e; }
          }
        }
      }
    }, KDepth<4>());
  }
  __syncthreads();
  // ---- order inside each bucket; exact e of the K kept tokens at their positions
  double* es = reinterpret_cast<double*>(hist);              // [K_TOPK_CAP]
  int* sid = reinterpret_cast<int*>(hist + 2 * K_TOPK_CAP);  // [K_TOPK_CAP]
  for (int p = tid; p < total; p += KT) {
    const KEntry me = list[p];
    int r = (int)me.ex;
    for (int o = (int)me.ex; o < total; ++o) {
      const KEntry ot = list[o];
      if (ot.ex != me.ex) break;
      r += (ot.key > me.key || (ot.key == me.key && ot.id < me.id)) ? 1 : 0;   // coder order: key, then lower id
    }
    if (r < K) {
      double x = (double)fmaxf(me.key, clamp_key);           // (double(x)/temp) - (double(max)/temp), :128-130
      if (!UNIT_TEMP) x = __ddiv_rn(x, P.temp);
      es[r] = ns_exp64_core(x - dm, c_exp_tab);
      sid[r] = me.id;
    }
  }
  __syncthreads();
  // ---- bin widths, prefix sums, overfill, selection: thread r holds the token of rank r
  const double ev = tid < K ? es[tid] : 0.0;
  double S = ev;                                             // sum of the kept e, fixed order (:146)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) S = S + __shfl_xor_sync(0xffffffffu, S, o);
  if (lane == 0) sc.red[warp] = (u64)__double_as_longlong(S);
  __syncthreads();
  S = __longlong_as_double((long long)sc.red[0]);
#pragma unroll
  for (int w = 1; w < KW; ++w) S = S + __longlong_as_double((long long)sc.red[w]);
  const double C = __ddiv_rn((double)R, S);
  const u64 q = tid < K ? (u64)__double2ll_rn(ev * C) : 0ull;   // :146-149
  u64 cum = q;                                               // inclusive prefix sums over the ranks (:150)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, cum, o); if (lane >= o) cum += t; }
  __syncthreads();                                           // red of the sum is consumed
  if (lane == 31) sc.red[warp] = cum;
  __syncthreads();
  u64 Q = 0;
  {
    u64 wo = 0;
#pragma unroll
    for (int w = 0; w < KW; ++w) { const u64 x = sc.red[w]; if (w < warp) wo += x; Q += x; }
    cum += wo;
  }
  u64* cums = reinterpret_cast<u64*>(list);                  // [K_TOPK_CAP]; the gathered list is no longer needed
  if (tid < K) cums[tid] = cum;
  __syncthreads();
  // overfill (:153-158): drop the ranks from the first prefix sum above the range on
  int kk = K;
  u64 slack;
  if (Q > R) {
    if (tid < K && cum > R && (tid == 0 || cums[tid - 1] <= R)) sc.res_idx = tid;
    __syncthreads();
    kk = sc.res_idx;
    slack = R - (kk > 0 ? cums[kk - 1] : 0ull);
    __syncthreads();
    if (tid == 0) sc.res_idx = K;
    __syncthreads();
  } else {
    slack = R - Q;
  }
  // bin of rank r: [cums[r-1] + slack, cums[r] + slack), rank 0 starts at 0 and absorbs the slack (:158)
  const u64 my_lo = (tid > 0 && tid < K) ? cums[tid - 1] + slack : 0ull;
  const u64 my_hi = cum + slack;
  if (MODE == MODE_ENC) {
    const u64 m_rel = window - lo;                           // next `precision` message bits (:168-171)
    if (tid < kk && my_lo <= m_rel && m_rel < my_hi) sc.res_idx = tid;   // :172 (empty bins never match)
    __syncthreads();
    const int r = sc.res_idx;
    if (tid == (r < kk ? r : 0)) {
      if (r >= kk && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW);   // cannot happen: the bins tile the range
      finish_encode(P, row, slot, sid[tid], lo + my_lo, lo + my_hi, (u64)K, Q, cursor, mlen);   // :175-176
    }
  } else {
    const bool tok_ok = tok >= 0 && tok < V;
    if (tid < kk && tok_ok && sid[tid] == tok) { sc.res_idx = tid; sc.res_found = 1; }
    __syncthreads();
    const bool in_range = sc.res_found != 0;
    const int r = in_range ? sc.res_idx : 0;                 // :342 / :347-348: unknown tokens are coded as rank 0
    if (tid == r) finish_decode(P, row, slot, in_range || !tok_ok, lo + my_lo, lo + my_hi, (u64)K, Q);
  }
}
