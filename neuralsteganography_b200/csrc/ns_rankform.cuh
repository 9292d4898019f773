// ns_rankform.cuh -- arithmetic-coder step when top-k binds (code_base/arithmetic.py:75: k = topk because more than topk
// tokens have p >= 1/range), 2 <= topk <= 512 < V: the reference's usual setting (run_single.py: topk 300).  sm_100a.
// Included by ns_coder.cu inside namespace nsr, after the shared definitions.
//
// Only the topk largest logits matter, so nothing of the row has to live on chip and no fp64 pass touches it: one
// 512-thread CTA per row, two per SM (38 KB of shared memory each), the row is swept twice from global memory --
//   sweep 1 (HBM, lines kept in L2)   fp32 online softmax: row maximum, estimate of sum exp, and every THREAD's own
//                                     maximum.  The topk-th largest thread maximum bounds the topk largest logits from
//                                     below (at least topk elements are that large): no sampling, no histogram of the row.
//   sweep 2 (L2)                      list the keys >= that bound (about 1.4 topk of them; one compare per element)
// then a count histogram of the listed keys groups them by bucket (order is local to a bucket), the bucket of position
// topk-1 is resolved exactly, and exp / sum / bin widths / prefix sums / overfill / search run on topk elements with
// one thread each -- the arithmetic of ns_fast.cuh's rank form and of the exact kernel, same integers.
// Rows in different phases share an SM, so the latency of one row's serial tail hides under the sweeps of the others.
// A row that is not certainly in rank form (not more than topk keys above the provisional cutoff + guard), a list
// overflow (flat rows, massive ties) or a dense boundary bucket goes to the exact kernel through slow_ws.

constexpr int RT = 512;                  // threads per CTA
constexpr int RW = RT / 32;
constexpr int R_NB = 2048;               // histogram buckets of the listed keys
constexpr int R_BPT = R_NB / RT;
constexpr int R_K_CAP = 512;             // topk the path holds (one thread per kept token)
constexpr int R_CAND_CAP = 2048;         // listed keys
constexpr int R_BND_CAP = 128;           // entries of the boundary bucket
constexpr int R_MIN_VOCAB = 4 * RT;      // every thread owns a chunk (its maximum enters the bound)
constexpr float R_BAND_EPS = 0.0009765625f;
static_assert(R_K_CAP <= RT, "one thread per kept token");

enum { R_WHY_EST = 1, R_WHY_RANK = 4, R_WHY_BUCKET = 6 };

struct RCand { float key; int id; };
struct RTop { uint32_t ebits; int id; uint32_t w; float key; };   // 16 B
struct RBnd { int id; float key; };

struct RScal {
  u64 red[3 * RW];
  int c_n, u_n, sel_bin, res_idx, res_found, pad;
  u64 sel_prefix;
  float tau; int tau_cnt;
};

__device__ __forceinline__ float r_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void r_hand_over(const ns_ac_params& P, int32_t* slow_ws, int row, int why) {
  const int s = atomicAdd(&slow_ws[0], 1);
  slow_ws[2 + s] = row;
  if (P.status) atomicOr(&P.status[row], NS_ST_EST_RETRY | (why << 8));
}
template <int N> struct RDepth { static constexpr int value = N; };

template <bool UNIT_TEMP, int MODE>
__global__ void __launch_bounds__(RT, 2) ac_rankform_kernel(const __grid_constant__ ns_ac_params P, int32_t* slow_ws) {
  __shared__ double tab[NS_EXP_N];
  __shared__ uint32_t hist[R_NB];                            // later: es[K] (double) + sid[K]
  __shared__ RCand cand[R_CAND_CAP];                         // later: cums[K] (u64)
  __shared__ RTop top[R_K_CAP];
  __shared__ RBnd bnd[R_BND_CAP];
  __shared__ uint32_t tmax[RT];
  __shared__ RScal sc;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, V = P.V, K = P.topk;
  long long tq = clock64();
  auto lap = [&](int k) { if (P.prof && tid == 0) { const long long t = clock64(); atomicAdd((unsigned long long*)&P.prof[k], (unsigned long long)(t - tq)); tq = t; } };
  int phase = P.phase ? (int)P.phase[row] : NS_PHASE_CODING;
  if (phase == NS_PHASE_DONE) return;
  if (MODE != MODE_ENC) phase = NS_PHASE_CODING;
  const int slot = P.ntok ? P.ntok[row] : 0;
  if (MODE == MODE_ENC && P.ntok && slot >= P.token_cap) {
    if (tid == 0) { if (P.phase) P.phase[row] = NS_PHASE_DONE; if (P.status) atomicOr(&P.status[row], NS_ST_TOKEN_OVERFLOW); }
    return;
  }
  if (MODE == MODE_DEC && P.ntok_total && slot >= P.ntok_total[row]) { if (tid == 0 && P.phase) P.phase[row] = NS_PHASE_DONE; return; }
  for (int i = tid; i < NS_EXP_N; i += RT) tab[i] = c_exp_tab[i];
  for (int i = tid; i < R_NB; i += RT) hist[i] = 0;
  if (tid == 0) { sc.c_n = 0; sc.u_n = 0; sc.sel_bin = -1; sc.sel_prefix = 0; sc.res_found = 0; }
  // the stream's scalars (their latency overlaps sweep 1)
  const u64 m_lo = P.lo[row], m_hi = P.hi[row];
  int m_cursor = 0, m_mlen = 0, m_tok = -1;
  u64 m_window = 0;
  if (MODE == MODE_ENC) {
    m_cursor = P.cursor[row]; m_mlen = P.msg_len[row];
    if (phase == NS_PHASE_CODING) m_window = ns_read_bits(P.msg + (size_t)row * P.msg_stride, m_cursor, m_mlen, P.precision);   // :168-171
  } else {
    m_tok = P.token_in[(size_t)row * P.token_stride + slot];
  }
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  const int W4 = (mis + V + 3) >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g - mis);
  const int mk0 = (P.mask_id[0] >= 0 && P.mask_id[0] < V) ? P.mask_id[0] : -8;
  const int mk1 = (P.mask_id[1] >= 0 && P.mask_id[1] < V) ? P.mask_id[1] : -8;
  float xmask[2];
  xmask[0] = mk0 >= 0 ? g[mk0] : -INFINITY;
  xmask[1] = mk1 >= 0 ? g[mk1] : -INFINITY;
  u64 pol_last, pol_first;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  auto ldg4 = [&](int c, bool last) -> float4 {
    float4 v;
#ifdef NS_NO_L2_HINT
    (void)last;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(g4 + c));
#else
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(g4 + c), "l"(last ? pol_first : pol_last));
#endif
    return v;
  };
  // ------------------------------------------------------------------ sweep 1: maximum, estimate, thread maxima
  const double temp = P.temp;
  const float c2 = (float)(1.4426950408889634 / temp);
  // Batches of U chunks: one maximum over the batch, one (rare) rescale, then 4U exp2.  The forbidden tokens are left in
  // here (no per-chunk test): their share of the estimate is taken out below, and a thread maximum that is a forbidden
  // token only loosens the bound on the top-k (the listed keys are counted again without them).
  float tm = -3.0e38f, ts = 0.f, ntc = 3.0e38f * c2;
  int ti = 0;
  {
    constexpr int U = 6;
    auto batch = [&](const float4* v, int c0, int n) {       // chunks c0, c0 + RT, ... (n of them, the rest -inf)
      float cm = -INFINITY;
#pragma unroll
      for (int u = 0; u < U; ++u) cm = fmaxf(cm, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
      if (cm > tm) {                                         // rare after the first batches
        ts *= r_ex2((tm - cm) * c2);
        tm = cm;
        ntc = -cm * c2;
        bool found = false;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int b = 4 * (c0 + u * RT) - mis;
          const float xs[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) if (!found && xs[e] == cm) { ti = b + e; found = true; }
        }
      }
      float part = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u)
        part += (r_ex2(fmaf(v[u].x, c2, ntc)) + r_ex2(fmaf(v[u].y, c2, ntc))) + (r_ex2(fmaf(v[u].z, c2, ntc)) + r_ex2(fmaf(v[u].w, c2, ntc)));
      ts += part;
      (void)n;
    };
    const float4 ninf4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int c = 1 + tid;
    for (; c + (U - 1) * RT < W4 - 1; c += U * RT) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ldg4(c + u * RT, false);
      batch(v, c, U);
    }
    if (c < W4 - 1 || tid == 0 || tid == 32) {               // the ragged end and the two edge chunks (element-wise, -inf outside the row)
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = (c + u * RT < W4 - 1) ? ldg4(c + u * RT, false) : ninf4;
      batch(v, c, U);
      if (tid == 0 || tid == 32) {
        const int ce = tid ? W4 - 1 : 0, b0 = 4 * ce - mis;
        float4 e[U];
#pragma unroll
        for (int u = 0; u < U; ++u) e[u] = ninf4;
        e[0].x = (b0 >= 0 && b0 < V) ? g[b0] : -INFINITY;
        e[0].y = (b0 + 1 >= 0 && b0 + 1 < V) ? g[b0 + 1] : -INFINITY;
        e[0].z = (b0 + 2 >= 0 && b0 + 2 < V) ? g[b0 + 2] : -INFINITY;
        e[0].w = (b0 + 3 >= 0 && b0 + 3 < V) ? g[b0 + 3] : -INFINITY;
        batch(e, ce, 1);
        if (tid == 0) {                                      // chunk 0 came last but holds the lowest ids: equal maxima go to it
          const float xs[4] = {e[0].w, e[0].z, e[0].y, e[0].x};
#pragma unroll
          for (int k = 0; k < 4; ++k) if (xs[k] == tm && b0 + 3 - k >= 0) ti = b0 + 3 - k;
        }
      }
    }
  }
  lap(0);                                                    // sweep 1
  float M, ssum;
  int top_id;
  {
    const uint32_t ok = ns_f32_orderable(tm + 0.0f);
    tmax[tid] = ok;
    const uint32_t wk = __reduce_max_sync(0xffffffffu, ok);
    const int wi = __reduce_min_sync(0xffffffffu, ok == wk ? ti : 0x7fffffff);
    const float wm = key_of_pack((u64)wk << 32);
    float wts = ts * r_ex2((tm - wm) * c2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wts += __shfl_xor_sync(0xffffffffu, wts, o);
    uint4* red4 = reinterpret_cast<uint4*>(sc.red);
    if (lane == 0) red4[warp] = make_uint4(wk, (uint32_t)wi, __float_as_uint(wts), 0u);
    __syncthreads();
    const uint4 pr = red4[lane < RW ? lane : 0];
    const uint32_t mk = __reduce_max_sync(0xffffffffu, pr.x);
    top_id = __reduce_min_sync(0xffffffffu, pr.x == mk ? (int)pr.y : 0x7fffffff);
    M = key_of_pack((u64)mk << 32);
    float part = lane < RW ? __uint_as_float(pr.z) * r_ex2((key_of_pack((u64)pr.x << 32) - M) * c2) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    ssum = part;
  }
  bool max_forbidden = false;
#pragma unroll
  for (int k = 0; k < 2; ++k) {                              // forbidden tokens (arithmetic.py:124-125): probability exactly 0
    const int id = k ? mk1 : mk0;
    if (id >= 0) {
      const float x = xmask[k];
      if (x > -INFINITY) ssum -= r_ex2((x - M) * c2);
      if (x >= M) max_forbidden = true;                      // rare: the exact kernel handles the row
    }
  }
  if (max_forbidden) { if (tid == 0) r_hand_over(P, slow_ws, row, R_WHY_EST); return; }
  if (MODE == MODE_ENC && phase == NS_PHASE_TAIL) {          // finish_sent tail: the rank-0 token (arithmetic.py:135-137)
    if (tid == 0) finish_tail(P, row, slot, top_id);
    return;
  }
  // ------------------------------------------------------------------ constants, the bound on the top-k
  const u64 lo = m_lo, R = m_hi - m_lo;                      // :140
  const double thr = __ddiv_rn(1.0, (double)R);              // :141
  const double Md = (double)M;
  const double dm = UNIT_TEMP ? Md : __ddiv_rn(Md, temp);
  const double theta_est = thr * (double)ssum;
  const float tf = (float)temp;
  const float key_th = fmaf(tf * 0.6931471805599453f, __log2f((float)theta_est), M);
  // top-k binds if more than topk tokens are above the cutoff even should the estimate be 2 % off
  const float kappa_r = key_th + tf * R_BAND_EPS + 0.02f * tf;
  const float clamp_key = (float)(Md - 700.0 * temp);
  if (!(ssum > 0.0f) || !(R >= 2) || !(kappa_r > clamp_key) || !(M - kappa_r > 0.0f)) {
    if (tid == 0) r_hand_over(P, slow_ws, row, R_WHY_EST);
    return;
  }
  // K-th largest thread maximum, from a count histogram of the RT maxima: every thread maximum in the buckets up to the one
  // holding position K-1 is at least tau, and there are at least K of them
  float tau;
  {
    uint32_t lowest = __reduce_min_sync(0xffffffffu, tmax[tid]);
    if (lane == 0) sc.red[2 * RW + warp] = lowest;           // (the first 2 RW words still hold the row reductions)
    __syncthreads();
    lowest = (uint32_t)sc.red[2 * RW + (lane < RW ? lane : 0)];
    lowest = __reduce_min_sync(0xffffffffu, lowest);
    const float tlo = key_of_pack((u64)lowest << 32);
    const float span = M - tlo;
    const float scale = span > 0.0f ? (float)(R_NB - 1) / span : 0.0f;
    const float mine = key_of_pack((u64)tmax[tid] << 32);
    int b = (int)((M - mine) * scale);
    b = b < 0 ? 0 : (b > R_NB - 1 ? R_NB - 1 : b);
    atomicAdd(&hist[b], 1u);
    __syncthreads();
    uint32_t hl[R_BPT], tsum = 0;
#pragma unroll
    for (int k = 0; k < R_BPT; ++k) { hl[k] = hist[tid * R_BPT + k]; tsum += hl[k]; }
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sc.red[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < RW; ++w) if (w < warp) woff += (uint32_t)sc.red[w];
    uint32_t excl = woff + inc - tsum;
#pragma unroll
    for (int k = 0; k < R_BPT; ++k) {
      if (hl[k] != 0 && excl <= (uint32_t)(K - 1) && (uint32_t)(K - 1) < excl + hl[k]) sc.sel_bin = tid * R_BPT + k;
      excl += hl[k];
      hist[tid * R_BPT + k] = 0;                             // clean for the listed keys
    }
    __syncthreads();
    const int bk = sc.sel_bin;
    uint32_t low_in = (b <= bk) ? tmax[tid] : 0xffffffffu;   // smallest thread maximum among the first buckets
    low_in = __reduce_min_sync(0xffffffffu, low_in);
    if (lane == 0) sc.red[RW + warp] = low_in;
    __syncthreads();
    low_in = (uint32_t)sc.red[RW + (lane < RW ? lane : 0)];
    low_in = __reduce_min_sync(0xffffffffu, low_in);
    tau = key_of_pack((u64)low_in << 32);
    if (tid == 0) { sc.sel_bin = -1; sc.sel_prefix = 0; }
  }
  lap(1);                                                    // reductions, constants, bound
  // ------------------------------------------------------------------ sweep 2: list the keys >= tau
  {
    auto visit = [&](const float4 v, int id) {
      const float hi4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
      if (hi4 >= tau) {                                      // rare
        const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = id + j;
          if (xs[j] >= tau && (unsigned)i < (unsigned)V && i != mk0 && i != mk1) {
            const int s2 = atomicAdd(&sc.c_n, 1);
            if (s2 < R_CAND_CAP) { cand[s2].key = xs[j] + 0.0f; cand[s2].id = i; }   // -0 -> +0: ties are ordered by id
          }
        }
      }
    };
    constexpr int U = 6;
    int c = 1 + tid;
    for (; c + (U - 1) * RT < W4 - 1; c += U * RT) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = ldg4(c + u * RT, true);
#pragma unroll
      for (int u = 0; u < U; ++u) visit(v[u], 4 * (c + u * RT) - mis);
    }
    for (; c < W4 - 1; c += RT) visit(ldg4(c, true), 4 * c - mis);
    if (tid == 0 || tid == 32) {
      const int ce = tid ? W4 - 1 : 0, b0 = 4 * ce - mis;
      float4 v;
      v.x = (b0 >= 0 && b0 < V) ? g[b0] : -INFINITY;
      v.y = (b0 + 1 >= 0 && b0 + 1 < V) ? g[b0 + 1] : -INFINITY;
      v.z = (b0 + 2 >= 0 && b0 + 2 < V) ? g[b0 + 2] : -INFINITY;
      v.w = (b0 + 3 >= 0 && b0 + 3 < V) ? g[b0 + 3] : -INFINITY;
      visit(v, b0);
    }
  }
  __syncthreads();
  lap(2);                                                    // sweep 2
  const int ncand = sc.c_n;
  if (ncand > R_CAND_CAP) { if (tid == 0) r_hand_over(P, slow_ws, row, R_WHY_BUCKET); return; }
  // rank form is certain iff more than K keys are >= kappa_r (every such key is listed when tau <= kappa_r, and when
  // tau > kappa_r all the listed keys are)
  {
    int above = 0;
    for (int j = tid; j < ncand; j += RT) above += cand[j].key >= kappa_r ? 1 : 0;
    above = __reduce_add_sync(0xffffffffu, above);
    if (lane == 0) sc.red[warp] = (u64)above;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int w = 0; w < RW; ++w) tot += (int)sc.red[w];
    if (tot <= K || ncand <= K) { if (tid == 0) r_hand_over(P, slow_ws, row, R_WHY_RANK); return; }
    __syncthreads();
  }
  lap(3);                                                    // rank-form check
  // ------------------------------------------------------------------ group the listed keys by bucket, find position K-1
  const float span2 = M - tau;
  const float rscale2 = span2 > 0.0f ? (float)R_NB / span2 : 0.0f, b_off2 = M * rscale2 + 8388608.0f;
  const float b_max = 8388608.0f + (float)(R_NB - 1);
  auto bucket_of = [&](float v) -> int {
    return __float_as_int(fmaxf(fminf(fmaf(-v, rscale2, b_off2), b_max), 8388608.0f)) & (R_NB - 1);
  };
  constexpr int CPT = R_CAND_CAP / RT;                       // listed keys per thread
  float ck[CPT]; int cid[CPT], cb[CPT];
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
    const int j = tid + u * RT;
    cid[u] = -1; ck[u] = 0.0f; cb[u] = 0;
    if (j < ncand) {
      cid[u] = cand[j].id; ck[u] = cand[j].key;
      cb[u] = bucket_of(ck[u]);
      atomicAdd(&hist[cb[u]], 1u);
    }
  }
  __syncthreads();
  uint32_t hloc[R_BPT];
  {
    uint32_t tsum = 0;
#pragma unroll
    for (int b = 0; b < R_BPT; ++b) { hloc[b] = hist[tid * R_BPT + b]; tsum += hloc[b]; }
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sc.red[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < RW; ++w) if (w < warp) woff += (uint32_t)sc.red[w];
    uint32_t excl = woff + inc - tsum, e2 = excl;
#pragma unroll
    for (int b = 0; b < R_BPT; ++b) {
      hist[tid * R_BPT + b] = hloc[b] | ((e2 < 0xffffu ? e2 : 0xffffu) << 16);   // count | exclusive prefix
      e2 += hloc[b];
    }
#pragma unroll
    for (int b = 0; b < R_BPT; ++b) {
      if (hloc[b] != 0 && excl <= (uint32_t)(K - 1) && (uint32_t)(K - 1) < excl + hloc[b]) { sc.sel_bin = tid * R_BPT + b; sc.sel_prefix = excl; }
      excl += hloc[b];
    }
    __syncthreads();
  }
  const int tb = sc.sel_bin;
  const int prefix = (int)sc.sel_prefix;                     // listed keys in the buckets before tb: all kept
#pragma unroll
  for (int u = 0; u < CPT; ++u) {
    if (cid[u] >= 0) {
      if (cb[u] < tb) {
        const uint32_t old = atomicSub(&hist[cb[u]], 1u);    // low 16 bits: slots still free in the bucket
        const uint32_t cnt_left = old & 0xffffu, ex = old >> 16;
        RTop e; e.ebits = ex; e.id = cid[u]; e.w = 0u; e.key = ck[u];
        top[ex + cnt_left - 1u] = e;
      } else if (cb[u] == tb) {
        const int s2 = atomicAdd(&sc.u_n, 1);
        if (s2 < R_BND_CAP) { bnd[s2].id = cid[u]; bnd[s2].key = ck[u]; }
      }
    }
  }
  __syncthreads();
  const int nbnd = sc.u_n;
  if (nbnd > R_BND_CAP) { if (tid == 0) r_hand_over(P, slow_ws, row, R_WHY_BUCKET); return; }
  auto before = [&](float ka, int ia, float kb, int ib) -> bool { return ka > kb || (ka == kb && ia < ib); };   // coder order
  {   // boundary bucket: its first K - prefix tokens in coder order complete the kept set, already in order
    const int need = K - prefix;
    if (tid < nbnd) {
      const float mk = bnd[tid].key;
      const int mi = bnd[tid].id;
      int r = 0;
      for (int o = 0; o < nbnd; ++o) r += before(bnd[o].key, bnd[o].id, mk, mi) ? 1 : 0;
      if (r < need) { RTop e; e.ebits = 0xffffffffu; e.id = mi; e.w = (uint32_t)(prefix + r); e.key = mk; top[prefix + r] = e; }
    }
  }
  __syncthreads();
  lap(4);                                                    // histogram of the listed keys, scatter, boundary bucket
  // ------------------------------------------------------------------ order inside each bucket, exp at the sorted positions
  double* es = reinterpret_cast<double*>(hist);              // [K]
  int* sid = reinterpret_cast<int*>(hist + 2 * R_K_CAP);     // [K]
  int my_r = -1, my_id = 0;
  float my_key = 0.0f;
  if (tid < K) {
    const RTop me = top[tid];
    my_id = me.id; my_key = me.key;
    if (me.ebits == 0xffffffffu) my_r = (int)me.w;           // boundary token: position known
    else {
      const int ex = (int)me.ebits;
      int r = ex;
      for (int o = ex; o < prefix && top[o].ebits == (uint32_t)ex; ++o) r += before(top[o].key, top[o].id, me.key, me.id) ? 1 : 0;
      my_r = r;
    }
  }
  __syncthreads();                                           // histogram words are free now
  auto a_of = [&](float key) -> double {                     // (double(x)/temp) - (double(max)/temp), :128-130
    double x = (double)fmaxf(key, clamp_key);
    if (!UNIT_TEMP) x = __ddiv_rn(x, temp);
    return x - dm;
  };
  if (tid < K) { es[my_r] = ns_exp64_core(a_of(my_key), tab); sid[my_r] = my_id; }
  __syncthreads();
  const double ev = tid < K ? es[tid] : 0.0;                 // thread r holds the token of rank r
  double S = ev;
  {                                                          // sum of the kept e, fixed order (:146)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) S = S + __shfl_xor_sync(0xffffffffu, S, o);
    if (lane == 0) sc.red[warp] = (u64)__double_as_longlong(S);
    __syncthreads();
    double rs = __longlong_as_double((long long)sc.red[0]);
#pragma unroll
    for (int w = 1; w < RW; ++w) rs = rs + __longlong_as_double((long long)sc.red[w]);
    S = rs;
    __syncthreads();
  }
  const double C = __ddiv_rn((double)R, S);
  const u64 q = tid < K ? (u64)__double2ll_rn(ev * C) : 0ull;   // :146-149
  u64 cum = q;                                               // inclusive prefix sums over the ranks (:150)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, cum, o); if (lane >= o) cum += t; }
  if (lane == 31) sc.red[warp] = cum;
  __syncthreads();
  u64 Q = 0;
  {
    u64 wo = 0;
#pragma unroll
    for (int w = 0; w < RW; ++w) { const u64 x = sc.red[w]; if (w < warp) wo += x; Q += x; }
    cum += wo;
  }
  u64* cums = reinterpret_cast<u64*>(cand);                  // [K]; the list is no longer needed
  if (tid < K) cums[tid] = cum;
  if (tid == 0) { sc.res_idx = K; sc.res_found = 0; }
  __syncthreads();
  lap(5);                                                    // order, exp, sum, widths, prefix sums
  // ------------------------------------------------------------------ overfill (:153-158), search, update
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
    const u64 m_rel = m_window - lo;                         // next `precision` message bits (:168-171)
    if (tid < kk && my_lo <= m_rel && m_rel < my_hi) sc.res_idx = tid;   // :172 (empty bins never match)
    __syncthreads();
    const int r = sc.res_idx;
    if (tid == (r < kk ? r : 0)) {
      if (r >= kk && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW);     // cannot happen: the bins tile the range
      finish_encode(P, row, slot, sid[tid], lo + my_lo, lo + my_hi, (u64)K, Q, m_cursor, m_mlen);   // :175-176
    }
  } else {
    int tok = m_tok;
    if (tok < 0 || tok >= V) tok = top_id;
    if (tid < kk && sid[tid] == tok) { sc.res_idx = tid; sc.res_found = 1; }
    __syncthreads();
    const bool in_range = sc.res_found != 0;
    const int r = in_range ? sc.res_idx : 0;                 // :342 / :347-348: unknown tokens are coded as rank 0
    if (tid == r) finish_decode(P, row, slot, in_range, lo + my_lo, lo + my_hi, (u64)K, Q);
  }
  lap(6);                                                    // overfill, search, update
  if (P.prof && tid == 0) atomicAdd((unsigned long long*)&P.prof[15], 1ull);
}
