// ns_codecs_stream.cuh -- rank codec and Huffman baseline without a resident row (sm_100a).
// Included by ns_codecs.cu inside its anonymous namespace, after the shared helpers.
//
// The row-resident kernels (rank_kernel, huffman_kernel: one 1024-thread CTA with the whole 201 KB row in shared memory)
// run one row per SM at a time: the HBM stream of a row and its compute never overlap, and neither do two rows.  These
// codecs only sweep the row two or three times with a handful of instructions per element, so the row does not need to
// live on chip: one 512-thread CTA per row with 17 KB of shared memory (2 CTAs per SM: 296 rows x 201 KB = 60 MB in
// flight, so that the later sweeps still find the row in L2 -- with 4 rows per SM they had been evicted), every sweep reads the logits
// from global memory -- the first from HBM, the later ones from L2 (the row was just read; 148 SMs x 4 rows x 201 KB is
// half of the 126 MB L2).  Rows in different phases share an SM, so HBM latency, L2 latency, barriers and the serial
// tree / resolve steps of one row hide under the sweeps of the others.
//   sweep 1  (HBM) extent: largest key (-0 folded, forbidden tokens at -1e10 on the fly: the logits are read-only), and with it
//            everything that needs no extent -- rank encode: the count histogram of the selection, its bucket range taken
//            from a SAMPLE of the row (one chunk per thread) because any monotone bucket function yields the same token;
//            rank decode: the tokens ranked before the observed one (decode is done after this one sweep)
//   sweep 2  (L2) Huffman: fp32 log_softmax normaliser + the few candidates of the top 2^b (bounded from below by the n-th
//            largest group maximum: no histogram); rank encode: gather the bucket of the wanted position, resolve it exactly
//   (rank)   the count of tokens with p > 0 is V whenever the row's smallest key lies above the fp64 underflow bound;
//            only otherwise one more sweep counts them
// Same integers as the row-resident kernels: selection by (key, lower id first).
// Serves: Huffman with bits_per_word <= 5 (n <= 32 group maxima), rank without top_p / min_prob (top_k clamps the count).

// CTAs per SM the register budget is cut for, measured (profiles/r2b_experiments.txt, 14): the kernels are bound by how
// many rows an SM has in flight, not by loads in flight per thread -- rank 4 (32 registers), Huffman 3 (40 registers),
// four chunk loads in flight on the HBM sweep (was: 2 CTAs, 8 loads)
#ifndef NSC_RANK_CTAS
#define NSC_RANK_CTAS 4
#endif
#ifndef NSC_HUF_CTAS
#define NSC_HUF_CTAS 3
#endif
#ifndef NSC_D1
#define NSC_D1 4
#endif
constexpr int CT = 512;                      // threads per CTA
constexpr int CW = CT / 32;
constexpr int C_NB = 2048;                   // histogram buckets
constexpr int C_BPT = C_NB / CT;
constexpr int C_LIST = 512;
template <int N> struct Depth { static constexpr int value = N; };   // chunk loads in flight per thread of a sweep

constexpr int C_GROUPS = 32;                 // lane groups whose maxima bound the Huffman selection
constexpr int C_GL = CT / C_GROUPS;          // lanes per group (8)
struct StreamScal {
  u64 red[CW];
  uint32_t wmax[C_GROUPS], wmin[CW];
  int list_count, sel_bin, res_idx, res_found;
  u64 sel_prefix;
};

// monotone bucket of a key without a conversion or a shift: (hi - k) * scale + 2^21 lands in [2^21, 2^22), where a float
// counts quarters -- its mantissa field is rint(4 (hi - k) scale), and that field with the two low bits masked off is
// the BYTE offset of bucket floor((hi - k) scale + 1/8) in the histogram.  Clamped to [0, NB) buckets.
constexpr float C_MAGIC = 2097152.0f;        // 2^21
__device__ __forceinline__ uint32_t stream_off(float k, float scale, float off) {
  return __float_as_uint(fmaxf(fminf(fmaf(-k, scale, off), C_MAGIC + (float)(C_NB - 1) + 0.75f), C_MAGIC)) & ((uint32_t)(C_NB - 1) << 2);
}

// Huffman tree of n <= 32 leaves by one warp: while all frequencies met so far are distinct, "pop the two smallest" has
// one answer and needs no heap -- two warp minima per merge.  Returns false at the first tie (heapq's sift order then
// decides, huffman.py:48-57): the caller falls back to the exact heapq replica.
__device__ __forceinline__ bool huf_build_warp(HufNode* nd, int n, int lane) {
  uint32_t f = lane < n ? __float_as_uint(nd[lane].freq) : 0x7f800000u;   // positive floats order like their bit patterns
  int id = lane;
  for (int step = 0; step + 1 < n; ++step) {
    const uint32_t m1 = __reduce_min_sync(0xffffffffu, f);
    const unsigned w1 = __ballot_sync(0xffffffffu, f == m1);
    if (__popc(w1) != 1) return false;
    const int l1 = __ffs(w1) - 1;
    const uint32_t g = lane == l1 ? 0x7f800000u : f;
    const uint32_t m2 = __reduce_min_sync(0xffffffffu, g);
    const unsigned w2 = __ballot_sync(0xffffffffu, g == m2);
    if (__popc(w2) != 1 || m2 == 0x7f800000u) return false;
    const int l2 = __ffs(w2) - 1;
    const int id1 = __shfl_sync(0xffffffffu, id, l1), id2 = __shfl_sync(0xffffffffu, id, l2);
    const float sum = __uint_as_float(m1) + __uint_as_float(m2);        // node1.freq + node2.freq in fp32
    const int nid = n + step;
    if (lane == 0) {
      nd[nid].freq = sum; nd[nid].left = id1; nd[nid].right = id2; nd[nid].parent = -1;   // left = first popped = bit 0
      nd[id1].parent = nid; nd[id2].parent = nid;
    }
    if (lane == l1) { f = __float_as_uint(sum); id = nid; }
    if (lane == l2) f = 0x7f800000u;
  }
  __syncwarp();
  return true;
}

static_assert(C_GROUPS == 32 && C_GL * C_GROUPS == CT && C_GL <= 32 && 32 % C_GL == 0, "one group maximum per lane of a warp");
template <int KIND>
__global__ void __launch_bounds__(CT, (KIND == K_RANK_ENC || KIND == K_RANK_DEC) ? NSC_RANK_CTAS : NSC_HUF_CTAS) codec_stream_kernel(ns_codec_params P) {
  constexpr bool RANK = KIND == K_RANK_ENC || KIND == K_RANK_DEC;
  constexpr bool DECODE = KIND == K_RANK_DEC || KIND == K_HUF_DEC;
  constexpr bool USE_MASK = !RANK;                           // the baselines forbid two tokens, the rank codec none
  __shared__ uint32_t hist[C_NB];                            // rank encode: count histogram; Huffman: ids / heap / nodes
  __shared__ ListEntry list[C_LIST];
  __shared__ StreamScal sc;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, V = P.V;
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  const int W4 = (mis + V + 3) >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g - mis);          // 16-byte aligned view
  const int mk0 = (USE_MASK && P.mask_id[0] >= 0 && P.mask_id[0] < V) ? P.mask_id[0] : -8;
  const int mk1 = (USE_MASK && P.mask_id[1] >= 0 && P.mask_id[1] < V) ? P.mask_id[1] : -8;
  const int mc0 = (mk0 + mis) >> 2, mc1 = (mk1 + mis) >> 2;
  // chunk c of the row as the codecs see it: -0 folded into +0, forbidden tokens at -1e10, -inf outside the row
  u64 pol_last, pol_first;                                   // L2 eviction policies for the row's lines
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  // `last`: nothing reads the row after this sweep (its lines may leave L2 first); earlier sweeps ask L2 to keep them
  auto ldg4 = [&](int c, bool last) -> float4 {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(g4 + c), "l"(last ? pol_first : pol_last));
    return v;
  };
  // rank encode: the sample chunk of this thread is requested before the stream's scalars are fetched (three dependent
  // loads), so that the row's first bytes are on their way while those resolve
  const int s_stride = (W4 - 2) / CT > 0 ? (W4 - 2) / CT : 1;
  const int s_cs = 1 + tid * s_stride;
  float4 s_v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (RANK && !DECODE && s_cs < W4 - 1) s_v = ldg4(s_cs, false);
  // ---- the stream's scalars
  const uint8_t phase = P.phase ? P.phase[row] : (uint8_t)NS_PHASE_CODING;
  const int slot = P.ntok ? P.ntok[row] : 0;
  if (phase == NS_PHASE_DONE) return;
  if (!DECODE && P.ntok && slot >= P.token_cap) {
    if (tid == 0) { if (P.phase) P.phase[row] = NS_PHASE_DONE; if (P.status) atomicOr(&P.status[row], NS_ST_TOKEN_OVERFLOW); }
    return;
  }
  const int total = (DECODE && P.ntok_total) ? P.ntok_total[row] : 0x7fffffff;
  if (DECODE && slot >= total) { if (tid == 0 && P.phase) P.phase[row] = NS_PHASE_DONE; return; }
  int cursor = 0, mlen = 0, tok = -1;
  uint32_t window = 0;
  if (!DECODE) {
    cursor = P.cursor[row]; mlen = P.msg_len[row];
    window = (uint32_t)ns_read_bits(P.msg + (size_t)row * P.msg_stride, cursor, mlen, 32);
  } else {
    tok = P.token_in[(size_t)row * P.token_stride + slot];
  }
  auto fold = [&](float4 v, int c) -> float4 {               // -0 -> +0; forbidden tokens (huffman_baseline.py:26-27)
    if (RANK) return v;                                      // the rank codec compares floats (-0 == +0) and folds where it packs a key
    v.x += 0.0f; v.y += 0.0f; v.z += 0.0f; v.w += 0.0f;
    if (USE_MASK && (c == mc0 || c == mc1)) {
      const int b = 4 * c - mis;
      float* f = reinterpret_cast<float*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) if (b + j == mk0 || b + j == mk1) f[j] = -1e10f;
    }
    return v;
  };
  // Sweep over the row: the interior chunks 1 .. W4-2 with C_UNROLL straight 128-bit loads in flight per thread (no
  // branch between them), then the two edge chunks element-wise by two threads.  body(v, first id, edge): `edge` chunks
  // carry -inf outside the row.
  auto sweep = [&](bool last, auto body, auto depth) {
    constexpr int U = decltype(depth)::value;                // loads in flight per thread
    int c = 1 + tid;
    for (; c + (U - 1) * CT < W4 - 1; c += U * CT) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ldg4(c + u * CT, last);
#pragma unroll
      for (int u = 0; u < U; ++u) body(fold(v[u], c + u * CT), 4 * (c + u * CT) - mis, false);
    }
    for (; c < W4 - 1; c += CT) body(fold(ldg4(c, last), c), 4 * c - mis, false);
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
  if (RANK && !DECODE) for (int i = tid; i < C_NB; i += CT) hist[i] = 0;
  if (tid == 0) sc.list_count = 0;
  // ---- rank encode: bucket range of the selection from a SAMPLE of the row (one chunk per thread, 8 KB), so that the
  // count histogram is filled by the same sweep that takes the exact extent.  Any monotone bucket function gives the
  // same token (the order inside a bucket is resolved exactly); a range that misses the true extent only clamps a few
  // keys into the end buckets.
  float scale = 0.0f, boff = C_MAGIC;
  if (RANK && !DECODE) {
    float sb = -INFINITY, sl = INFINITY;
    const int cs = s_cs;
    if (cs < W4 - 1) {
      const float4 v = fold(s_v, cs);
      sb = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      sl = fminf(fminf(v.x > -1e9f ? v.x : INFINITY, v.y > -1e9f ? v.y : INFINITY),
                 fminf(v.z > -1e9f ? v.z : INFINITY, v.w > -1e9f ? v.w : INFINITY));
    }
    const uint32_t a = __reduce_max_sync(0xffffffffu, ns_f32_orderable(sb));
    const uint32_t b = __reduce_min_sync(0xffffffffu, ns_f32_orderable(sl));
    if (lane == 0) sc.red[warp] = ((u64)a << 32) | (u64)b;
    __syncthreads();
    uint32_t ga = 0, gb = 0xffffffffu;
#pragma unroll
    for (int w = 0; w < CW; ++w) { const u64 r = sc.red[w]; ga = max(ga, (uint32_t)(r >> 32)); gb = min(gb, (uint32_t)r); }
    float smax = key_of_pack((u64)ga << 32), smin = key_of_pack((u64)gb << 32);
    if (!(smax > -3.0e38f) || !(smax < 3.0e38f) || !(smin > -3.0e38f) || !(smin < 3.0e38f)) { smax = 1.0f; smin = -1.0f; }
    float span = smax - smin;
    if (!(span > 0.0f)) span = 1.0f;
    const float hi_p = smax + 0.25f * span, lo_p = smin - 0.25f * span;   // the sample misses the tails: widen
    scale = (float)C_NB / (hi_p - lo_p);
    boff = hi_p * scale + C_MAGIC;
    if (!(scale > 0.0f) || !(scale < 3.0e38f) || !(fabsf(boff) < 3.0e38f)) { scale = 0.0f; boff = C_MAGIC; }
    __syncthreads();                                         // sc.red is reused below
  }
  // ---- sweep 1 (HBM): the extent; rank: with it everything that does not need the extent -- the count histogram
  // (encode), the tokens ranked before the observed one (decode)
  float bk = -INFINITY, ak = INFINITY;
  float tkey = INFINITY;
  if (RANK && DECODE) {
    if (tok >= 0 && tok < V) tkey = g[tok] + 0.0f;           // else: an unknown token counts as rank 0 (nothing ranks before +inf)
    else tok = -1;
  }
  uint32_t cnt = 0, before = 0;
  const uint32_t hb = (uint32_t)__cvta_generic_to_shared(hist);
  sweep(RANK && DECODE, [&](const float4 v, int id, bool edge) {
    bk = fmaxf(bk, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    if (RANK) {
      const float xs[4] = {v.x, v.y, v.z, v.w};
      if (!edge) {
        ak = fminf(ak, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (DECODE) before += (xs[j] > tkey || (xs[j] == tkey && id + j < tok)) ? 1u : 0u;   // coder order: key, then lower id
          else asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hb + stream_off(xs[j], scale, boff)), "r"(1u) : "memory");
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if ((unsigned)(id + j) >= (unsigned)V) continue;   // padding of the edge chunks is not a token
          ak = fminf(ak, xs[j]);
          if (DECODE) before += (xs[j] > tkey || (xs[j] == tkey && id + j < tok)) ? 1u : 0u;
          else asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hb + stream_off(xs[j], scale, boff)), "r"(1u) : "memory");
        }
      }
    }
  }, Depth<NSC_D1>());   // from HBM: more bytes in flight
  {
    const unsigned gmask = ((1u << C_GL) - 1u) << (C_GL * (lane / C_GL));       // this lane's group of C_GL lanes
    const uint32_t a = __reduce_max_sync(gmask, ns_f32_orderable(bk));
    const uint32_t c = RANK ? __reduce_min_sync(0xffffffffu, ns_f32_orderable(ak)) : 0u;
    if (lane % C_GL == 0) sc.wmax[tid / C_GL] = a;
    if (lane == 0) sc.wmin[warp] = c;
  }
  __syncthreads();
  const uint32_t gmax = __reduce_max_sync(0xffffffffu, sc.wmax[lane]);
  const uint32_t gall = __reduce_min_sync(0xffffffffu, sc.wmin[lane & (CW - 1)]);
  const float mx = key_of_pack((u64)gmax << 32), amin = key_of_pack((u64)gall << 32);   // amin: smallest key of the row (rank)

  auto block_sum_u = [&](u64 v) -> u64 {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sc.red[warp] = v;
    __syncthreads();
    u64 r = 0;
#pragma unroll
    for (int w = 0; w < CW; ++w) r += sc.red[w];
    return r;
  };

  if (RANK) {
    // tokens with p > 0 (codec/arithmetic.py:372): the fp64 softmax underflows below exp(-745).  Every token of an
    // ordinary row has p > 0 (its smallest key lies above the bound): the count is V; otherwise one more sweep counts --
    // a float compare except within a hair of the bound, where the fp64 expression decides.
    const double dm = (double)mx / P.temp;
    u64 n_pos = P.topk > 0 ? (u64)P.topk : ~0ull;            // quality.py:76-81
    const float kc = (float)((dm - 745.0) * P.temp);
    const float margin = fmaxf(1e-3f, fabsf(kc) * 1e-5f);
    const float k_yes = kc + margin, k_no = kc - margin;
    if (amin > k_yes) cnt = tid == 0 ? (uint32_t)V : 0u;
    else {
      sweep(DECODE, [&](const float4 v, int id, bool) {
        const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float k = xs[j];
          if ((unsigned)(id + j) >= (unsigned)V) continue;
          if (k > k_yes) cnt += 1u;
          else if (k >= k_no) cnt += (((double)k / P.temp - dm) >= -745.0) ? 1u : 0u;
        }
      }, Depth<4>());
    }
    const u64 both = block_sum_u(((u64)before << 40) | (u64)cnt);
    const u64 total_pos = both & 0xffffffffffull;
    if (total_pos < n_pos) n_pos = total_pos;
    const int capacity = n_pos ? 63 - __clzll((long long)n_pos) : 0;      // floor(log2(n_pos)), :379
    if (capacity <= 0) {                                                  // ArithmeticRangeError :149
      if (tid == 0 && P.status) atomicOr(&P.status[row], NS_ST_OUT_OF_RANGE);
      if (tid == 0 && P.phase) P.phase[row] = NS_PHASE_DONE;
      return;
    }
    if (!DECODE) {
      const u64 index = (u64)(window >> (32 - capacity));                 // :153-157
      u64 prefix = 0;
      for (int attempt = 0; attempt < 2; ++attempt) {
        // bucket of position `index`: scan of the count histogram
        uint32_t hl[C_BPT], tsum = 0;
  #pragma unroll
        for (int b = 0; b < C_BPT; ++b) { hl[b] = hist[tid * C_BPT + b]; tsum += hl[b]; }
        uint32_t inc = tsum;
  #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        __syncthreads();
        if (lane == 31) sc.red[warp] = inc;
        if (tid == 0) { sc.sel_bin = -1; sc.sel_prefix = 0; sc.res_found = 0; }
        __syncthreads();
        uint32_t woff = 0;
  #pragma unroll
        for (int w = 0; w < CW; ++w) if (w < warp) woff += (uint32_t)sc.red[w];
        uint32_t excl = woff + inc - tsum;
  #pragma unroll
        for (int b = 0; b < C_BPT; ++b) {
          if (hl[b] != 0 && excl <= index && index < excl + hl[b]) { sc.sel_bin = tid * C_BPT + b; sc.sel_prefix = excl; }
          excl += hl[b];
        }
        __syncthreads();
        const int tb = sc.sel_bin;
        prefix = sc.sel_prefix;
        // sweep 3 (L2): gather that bucket
        if (tb >= 0) {
          // keys of bucket tb lie within one bucket width of its centre: a two-instruction test per element, the exact
          // bucket function only for those that pass
          // bucket tb holds the keys with (hi - k) scale in [tb - 1/8, tb + 7/8) (hi = the rounded offset of the bucket
          // function); the end buckets also hold everything the range clamps.  One test per chunk, the exact bucket
          // function only for the chunks that pass.
          const float hi_p = (boff - C_MAGIC) / (scale > 0.0f ? scale : 1.0f);
          const bool endb = tb == 0 || tb == C_NB - 1;
          const float centre = scale > 0.0f ? hi_p - ((float)tb + 0.375f) / scale : mx, reach = (scale > 0.0f && !endb) ? 1.0f / scale : INFINITY;
          const uint32_t tboff = (uint32_t)tb << 2;
          sweep(true, [&](const float4 v, int id, bool) {
            const float near = fminf(fminf(fabsf(v.x - centre), fabsf(v.y - centre)), fminf(fabsf(v.z - centre), fabsf(v.w - centre)));
            if (near <= reach) {
              const float xs[4] = {v.x, v.y, v.z, v.w};
  #pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (stream_off(xs[j], scale, boff) == tboff && (unsigned)(id + j) < (unsigned)V) {
                  const int s2 = atomicAdd(&sc.list_count, 1);
                  if (s2 < C_LIST) { list[s2].pack = pack_of(xs[j] + 0.0f, id + j); list[s2].w = 1; }
                }
              }
            }
          }, Depth<4>());
        }
        __syncthreads();
        if (sc.list_count <= C_LIST || attempt == 1) break;
        // The sampled range put more keys into one bucket than the list holds (a row whose tail lies far outside the
        // sample): once more with the exact extent, as many sweeps as it takes -- the row is in L2.
        float lk = INFINITY;
        sweep(false, [&](const float4 v, int, bool) {
          lk = fminf(lk, fminf(fminf(v.x > -1e9f ? v.x : INFINITY, v.y > -1e9f ? v.y : INFINITY),
                               fminf(v.z > -1e9f ? v.z : INFINITY, v.w > -1e9f ? v.w : INFINITY)));
        }, Depth<4>());
        const uint32_t wl = __reduce_min_sync(0xffffffffu, ns_f32_orderable(lk));
        __syncthreads();
        if (lane == 0) sc.wmin[warp] = wl;
        for (int i = tid; i < C_NB; i += CT) hist[i] = 0;
        if (tid == 0) sc.list_count = 0;
        __syncthreads();
        const float mn = key_of_pack((u64)__reduce_min_sync(0xffffffffu, sc.wmin[lane & (CW - 1)]) << 32);
        const float span = mx - mn;
        scale = (span > 0.0f && span < 3.0e38f) ? (float)C_NB / span : 0.0f;
        boff = mx * scale + C_MAGIC;
        if (!(fabsf(boff) < 3.0e38f)) { scale = 0.0f; boff = C_MAGIC; }
        sweep(false, [&](const float4 v, int id, bool edge) {
          const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (!edge || (unsigned)(id + j) < (unsigned)V)
              asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hb + stream_off(xs[j], scale, boff)), "r"(1u) : "memory");
        }, Depth<4>());
        __syncthreads();
      }
      int nl = sc.list_count;
      if (nl > C_LIST) { nl = C_LIST; if (tid == 0 && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW); }
      if (tid < nl) {                                        // exact position inside the bucket
        const u64 pc = list[tid].pack;
        u64 pos = prefix;
        for (int o = 0; o < nl; ++o) pos += list[o].pack > pc ? 1 : 0;
        if (pos == index) { sc.res_idx = id_of_pack(pc); sc.res_found = 1; }
      }
      __syncthreads();
      int token = sc.res_found ? sc.res_idx : -1;
      if (token < 0) {                                       // cannot happen (index < 2^capacity <= n_pos): lowest id of the maximum
        int ti = 0x7fffffff;
        sweep(true, [&](const float4 v, int id, bool) {
          if (v.w == mx) ti = min(ti, id + 3);
          if (v.z == mx) ti = min(ti, id + 2);
          if (v.y == mx) ti = min(ti, id + 1);
          if (v.x == mx) ti = min(ti, id);
        }, Depth<4>());
        ti = __reduce_min_sync(0xffffffffu, ti);
        __syncthreads();
        if (lane == 0) sc.red[warp] = (u64)(uint32_t)ti;
        __syncthreads();
        for (int w = 0; w < CW; ++w) ti = min(ti, (int)sc.red[w]);
        token = ti;
      }
      int consumed = mlen - cursor;
      if (consumed > capacity) consumed = capacity;                       // :155
      if (tid == 0) emit_token(P, row, slot, token, consumed);
    } else {
      const u64 rank = both >> 40;                                        // ranked_tokens.index(token), :211
      if (tid == 0) {
        if ((tok < 0 || rank >= (1ull << capacity)) && P.status) atomicOr(&P.status[row], NS_ST_OUT_OF_RANGE);   // :212-213
        const int tot_bits = P.total_bits ? P.total_bits[row] : 0x7fffffff;
        int takeb = tot_bits - P.out_len[row];
        if (takeb > capacity) takeb = capacity;
        if (takeb < 0) takeb = 0;
        emit_bits(P, row, slot, (rank & ((1ull << capacity) - 1ull)) >> (capacity - takeb), takeb);   // :215-216
      }
    }
  } else {
    // ---- Huffman: the top n = 2^b tokens.  The n-th largest warp maximum bounds them from below (there are at least n
    // elements that large), so the candidates are the few keys >= that bound -- no histogram, no second selection sweep.
    int n = 1 << P.param;
    if (n > V) n = V;
    float tau;
    {
      const uint32_t mine = sc.wmax[lane];
      int r = 0;
#pragma unroll
      for (int j = 0; j < C_GROUPS; ++j) {
        const uint32_t o = __shfl_sync(0xffffffffu, mine, j);
        r += (o > mine || (o == mine && j < lane)) ? 1 : 0;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, r == n - 1);
      const uint32_t bound = __shfl_sync(0xffffffffu, mine, hit ? __ffs(hit) - 1 : 0);
      tau = (hit && bound) ? key_of_pack((u64)bound << 32) : -INFINITY;
    }
    // log_softmax normaliser (:32): fp32 exp per element, four at a time summed in fp32, the partial sums in double
    // (the reference sums in fp32; anything within its ~1e-7 is the same normaliser)
    double acc0 = 0.0, acc1 = 0.0;
    int flip = 0;
    sweep(true, [&](const float4 v, int id, bool) {
      const float s4 = (expf(v.x - mx) + expf(v.y - mx)) + (expf(v.z - mx) + expf(v.w - mx));
      if (flip) acc1 += (double)s4; else acc0 += (double)s4;
      flip ^= 1;
      const float hi = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      if (hi >= tau) {                                       // rare
        const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (xs[j] >= tau && id + j >= 0 && id + j < V) {
            const int s2 = atomicAdd(&sc.list_count, 1);
            if (s2 < C_LIST) { list[s2].pack = pack_of(xs[j], id + j); list[s2].w = 0; }
          }
        }
      }
    }, Depth<4>());
    double sum = acc0 + acc1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) sc.red[warp] = (u64)__double_as_longlong(sum);
    __syncthreads();
    int nc = sc.list_count;
    if (nc > C_LIST) { nc = C_LIST; if (tid == 0 && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW); }
    int* ids = reinterpret_cast<int*>(hist);                 // [32] ids by rank
    int* heap = ids + 32;                                    // [32]
    HufNode* nd = reinterpret_cast<HufNode*>(hist + 64);     // [63] nodes
    if (tid < nc) {                                          // rank among the candidates; the first n are the kept set
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < CW; ++w) tot += __longlong_as_double((long long)sc.red[w]);
      const float lse = (float)log(tot);
      const u64 pc = list[tid].pack;
      int r = 0;
      for (int o = 0; o < nc; ++o) r += list[o].pack > pc ? 1 : 0;
      if (r < n) {
        ids[r] = id_of_pack(pc);
        nd[r].freq = expf((key_of_pack(pc) - mx) - lse);     // probs = exp(log_probs), :33
        nd[r].left = -1; nd[r].right = -1; nd[r].parent = -1;
      }
    }
    __syncthreads();
    if (warp == 0) {
      const bool fast = huf_build_warp(nd, n, lane);
      int root = 2 * n - 2;
      if (!fast && lane == 0) {                              // equal frequencies: heapq's sift order decides (huffman.py:48-57)
        for (int i = 0; i < n; ++i) { nd[i].left = -1; nd[i].right = -1; nd[i].parent = -1; }
        int len = 0;
        for (int i = 0; i < n; ++i) { heap[len] = i; ++len; huf_siftdown(heap, nd, 0, len - 1); }   // make_heap_from_array
        int nxt = n;
        while (len > 1) {
          const int n1 = huf_pop(heap, nd, &len);
          const int n2 = huf_pop(heap, nd, &len);
          nd[nxt].freq = nd[n1].freq + nd[n2].freq;
          nd[nxt].left = n1; nd[nxt].right = n2; nd[nxt].parent = -1;
          nd[n1].parent = nxt; nd[n2].parent = nxt;
          heap[len] = nxt; ++len; huf_siftdown(heap, nd, 0, len - 1);
          ++nxt;
        }
        root = heap[0];
      }
      root = __shfl_sync(0xffffffffu, root, 0);
      if (lane == 0) {
        if (!DECODE) {
          int node = root, used = 0;
          while (nd[node].left >= 0) {                       // :47-52, exhausted message reads as 0
            const int bit = (int)((window >> (31 - used)) & 1u);
            node = bit ? nd[node].right : nd[node].left;
            ++used;
          }
          emit_token(P, row, slot, ids[node], used);
        } else {
          int leaf = -1;
          for (int r = 0; r < n; ++r) if (ids[r] == tok) leaf = r;
          if (leaf < 0) { leaf = 0; if (P.status) atomicOr(&P.status[row], NS_ST_OUT_OF_RANGE); }   // :149
          u64 code = 0;
          int depth = 0;
          for (int node = leaf; nd[node].parent >= 0; node = nd[node].parent) {
            const int par = nd[node].parent;
            code |= (u64)(nd[par].right == node ? 1 : 0) << depth;
            ++depth;
          }
          emit_bits(P, row, slot, code, depth);
        }
      }
    }
  }
}
