// ns_fast.cuh -- single-exp-pass arithmetic-coder step (the throughput path), sm_100a.
// Included by ns_coder.cu after the shared definitions (u64, pack_of, ListEntry, finish_*).
//
// Persistent CTAs (one per SM, 512 threads), each looping over rows.  Per row:
//   L   128-bit global loads -> shared words[] ; fused fp32 online softmax estimate
//       (row max, its lowest id, sum of exp) ; L2 prefetch of the CTA's next row
//   P1  ONE fp64 exp per element (10 fp64 ops): exact sum of all e_i in a fixed order, exact sum of
//       the provisionally-cut ones, elements within 2^-10 of the provisional cutoff go to a small
//       list with their exact e ; the word is overwritten in place by trunc_fp32(e_i) (0 if not kept)
//   FIX exact normaliser -> the provisional cutoff is verified, list elements classified exactly,
//       S_kept and C = range / S_kept exact
//   P2  q_i = rint(e_i * C) from the truncated e_i with a rigorous interval test (2 fp64 FMAs);
//       the few undecidable ones are redone exactly from the original logit (L2 hit) ;
//       integer mass histogram over 2048 monotone buckets of the fp32 bit pattern ; total mass
//   SEL/UPD as in the exact kernel (bucket prefix -> collect -> exact order by original logit).
// Anything unusual (top-k smaller than the cutoff set, list overflow, estimate outside its guard
// band) queues the row in slow_ws; the exact multi-pass kernel then redoes it.  The decision only
// depends on the row and its range, never on encode/decode, so both directions take the same path.

constexpr int FT = 512;              // threads per CTA
constexpr int FW = FT / 32;
constexpr int F_NB = 2048;           // histogram buckets (u32 masses: precision <= 31)
constexpr int F_BAND_CAP = 128;
constexpr int F_U_CAP = 256;
constexpr int F_C_CAP = 256;
constexpr float F_BAND_EPS = 0.0009765625f;   // 2^-10 half-width (in log units) of the exact-list band
constexpr uint32_t F_TOP = 0x3F800000u;       // bit pattern of 1.0f = e of the row maximum

struct BandEntry { int id; int kept; double e; };

struct FScal {
  u64 red[FW];
  float M; int top_id; float sum32; int remax;
  int band_n; int u_n; int c_n; int bail;
  u64 band_cut_int;
  int sel_bin; u64 sel_prefix;
  int res_idx; u64 res_before; u64 res_w; int res_found;
  // row constants (written by thread 0, read by everybody)
  double dm, thr; float kappa_lo, kappa_hi, clamp_key; int band_E;
};
static_assert(sizeof(FScal) <= 1024, "FScal too large");
static_assert(NS_EXP_N * 8 + F_NB * 4 + F_BAND_CAP * 16 + F_U_CAP * 4 + F_C_CAP * 16 + 1024 <= FIXED_BYTES, "fast smem layout");

__device__ __forceinline__ float f_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 f_ldg4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void f_prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

// queue a row for the exact kernel: slow_ws = {count, done, rows...}
__device__ __forceinline__ void hand_over(const ns_ac_params& P, int32_t* slow_ws, int row) {
  const int s = atomicAdd(&slow_ws[0], 1);
  slow_ws[2 + s] = row;
  if (P.status) atomicOr(&P.status[row], NS_ST_EST_RETRY);   // informational: row took the exact path
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
__device__ __forceinline__ double f_sum_d(double v, u64* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = (u64)__double_as_longlong(v);
  __syncthreads();
  double r = __longlong_as_double((long long)scratch[0]);
#pragma unroll
  for (int w = 1; w < FW; ++w) r = r + __longlong_as_double((long long)scratch[w]);
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

// online softmax step of one lane-private accumulator: (tm, ts, ti) <- element (x, id)
#define F_ONLINE(x, id, tm, ts, ti)                               \
  do {                                                            \
    if ((x) > (tm)) { (ts) *= f_ex2(((tm) - (x)) * c2); (tm) = (x); (ti) = (id); } \
    (ts) += f_ex2(((x) - (tm)) * c2);                             \
  } while (0)

template <bool UNIT_TEMP, int MODE>
__global__ void __launch_bounds__(FT, 1) ac_fast_kernel(ns_ac_params P, int32_t* slow_ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* tab = reinterpret_cast<double*>(smem_raw);
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + NS_EXP_N * 8);
  BandEntry* band = reinterpret_cast<BandEntry*>(smem_raw + NS_EXP_N * 8 + F_NB * 4);
  int* ulist = reinterpret_cast<int*>(smem_raw + NS_EXP_N * 8 + F_NB * 4 + F_BAND_CAP * 16);
  ListEntry* clist = reinterpret_cast<ListEntry*>(smem_raw + NS_EXP_N * 8 + F_NB * 4 + F_BAND_CAP * 16 + F_U_CAP * 4);
  FScal* sc = reinterpret_cast<FScal*>(smem_raw + NS_EXP_N * 8 + F_NB * 4 + F_BAND_CAP * 16 + F_U_CAP * 4 + F_C_CAP * 16);
  float* words = reinterpret_cast<float*>(smem_raw + FIXED_BYTES);   // element id lives at words[id + mis]
  float4* w4 = reinterpret_cast<float4*>(words);

  const int tid = threadIdx.x;
  const int V = P.V;
  const double temp = P.temp;
  const float c2 = (float)(1.4426950408889634 / temp);     // log2(e)/temp for the fp32 estimate
  const double magic = 6755399441055744.0;                 // 1.5 * 2^52

  for (int i = tid; i < NS_EXP_N; i += FT) tab[i] = c_exp_tab[i];

  for (int row = blockIdx.x; row < P.B; row += gridDim.x) {
    __syncthreads();                                       // previous row is finished with shared memory
    uint8_t phase = P.phase ? P.phase[row] : (uint8_t)NS_PHASE_CODING;
    if (phase == NS_PHASE_DONE) continue;
    if (MODE != MODE_ENC) phase = NS_PHASE_CODING;
    const int slot = P.ntok ? P.ntok[row] : 0;
    if (MODE == MODE_ENC && P.ntok && slot >= P.token_cap) {
      if (tid == 0) {
        if (P.phase) P.phase[row] = NS_PHASE_DONE;
        if (P.status) atomicOr(&P.status[row], NS_ST_TOKEN_OVERFLOW);
      }
      continue;
    }
    if (MODE == MODE_DEC && P.ntok_total && slot >= P.ntok_total[row]) {
      if (tid == 0 && P.phase) P.phase[row] = NS_PHASE_DONE;
      continue;
    }

    // ------------------------------------------------------------------ L: load + estimate
    const float* g = P.logits + (size_t)row * (size_t)P.ld;
    const int mis = (int)(((uintptr_t)g & 15u) >> 2);
    const int W4 = (mis + V + 3) >> 2;                     // float4 chunks of the padded row
    const float4* g4 = reinterpret_cast<const float4*>(g - mis);
    for (int i = tid; i < F_NB / 4; i += FT) reinterpret_cast<uint4*>(hist)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { sc->band_n = 0; sc->u_n = 0; sc->c_n = 0; sc->bail = 0; sc->band_cut_int = 0; sc->remax = 0; }
    float tm0 = -3.0e38f, tm1 = -3.0e38f, tm2 = -3.0e38f, tm3 = -3.0e38f;
    float ts0 = 0.f, ts1 = 0.f, ts2 = 0.f, ts3 = 0.f;
    int ti0 = 0, ti1 = 0, ti2 = 0, ti3 = 0;
    for (int c = tid; c < W4; c += FT) {
      float4 v;
      const int b = 4 * c - mis;
      if (c > 0 && c < W4 - 1) {
        v = f_ldg4(g4 + c);
      } else {
        v.x = (b >= 0 && b < V) ? g[b] : -INFINITY;
        v.y = (b + 1 >= 0 && b + 1 < V) ? g[b + 1] : -INFINITY;
        v.z = (b + 2 >= 0 && b + 2 < V) ? g[b + 2] : -INFINITY;
        v.w = (b + 3 >= 0 && b + 3 < V) ? g[b + 3] : -INFINITY;
      }
      v.x += 0.0f; v.y += 0.0f; v.z += 0.0f; v.w += 0.0f;  // -0 -> +0 (equal logits tie by id)
      w4[c] = v;
      F_ONLINE(v.x, b, tm0, ts0, ti0);
      F_ONLINE(v.y, b + 1, tm1, ts1, ti1);
      F_ONLINE(v.z, b + 2, tm2, ts2, ti2);
      F_ONLINE(v.w, b + 3, tm3, ts3, ti3);
    }
    {   // prefetch this CTA's next row into L2 while this one is processed
      const int nrow = row + gridDim.x;
      if (nrow < P.B) {
        const char* np = reinterpret_cast<const char*>(P.logits + (size_t)nrow * (size_t)P.ld);
        const int nbytes = V * 4;
        for (int off = tid * 128; off < nbytes; off += FT * 128) f_prefetch_l2(np + off);
      }
    }
    float Mt = fmaxf(fmaxf(tm0, tm1), fmaxf(tm2, tm3));
    u64 pk = 0;
    {
      u64 p0 = pack_of(tm0, ti0), p1 = pack_of(tm1, ti1), p2 = pack_of(tm2, ti2), p3 = pack_of(tm3, ti3);
      pk = p0 > p1 ? p0 : p1;
      u64 pq = p2 > p3 ? p2 : p3;
      pk = pk > pq ? pk : pq;
    }
    float tst = ts0 * f_ex2((tm0 - Mt) * c2) + ts1 * f_ex2((tm1 - Mt) * c2) + ts2 * f_ex2((tm2 - Mt) * c2) +
                ts3 * f_ex2((tm3 - Mt) * c2);
    const u64 pmax = f_reduce_u(pk, OpMaxU(), sc->red);
    float M = key_of_pack(pmax);
    int top_id = id_of_pack(pmax);
    float ssum = f_sum_f(tst * f_ex2((Mt - M) * c2), sc->red);
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
        const float4 v = w4[c];
        const int b = 4 * c - mis;
        u64 p;
        p = pack_of(v.x, b); pm = p > pm ? p : pm;
        p = pack_of(v.y, b + 1); pm = p > pm ? p : pm;
        p = pack_of(v.z, b + 2); pm = p > pm ? p : pm;
        p = pack_of(v.w, b + 3); pm = p > pm ? p : pm;
      }
      pm = f_reduce_u(pm, OpMaxU(), sc->red);
      M = key_of_pack(pm);
      top_id = id_of_pack(pm);
      float s = 0.f;
      for (int c = tid; c < W4; c += FT) {
        const float4 v = w4[c];
        s += f_ex2((v.x - M) * c2) + f_ex2((v.y - M) * c2) + f_ex2((v.z - M) * c2) + f_ex2((v.w - M) * c2);
      }
      s = f_sum_f(s, sc->red);
      if (tid == 0) sc->sum32 = s;
      __syncthreads();
    }
    ssum = sc->sum32;

    if (MODE == MODE_ENC && phase == NS_PHASE_TAIL) {
      if (tid == 0) finish_tail(P, row, slot, top_id);
      continue;
    }

    // ------------------------------------------------------------------ row constants
    const u64 lo = P.lo[row], hi = P.hi[row];
    const u64 R = hi - lo;                                   // arithmetic.py:140
    const double thr = __ddiv_rn(1.0, (double)R);            // :141
    const double Md = (double)M;
    const double dm = UNIT_TEMP ? Md : __ddiv_rn(Md, temp);
    if (tid == 0) {
      const double theta_est = thr * (double)ssum;
      int bail = !(ssum > 0.0f) || !(R >= 2);
      const double a_th = log(theta_est);
      const double key_th = Md + temp * a_th;
      sc->kappa_hi = (float)(key_th + temp * (double)F_BAND_EPS);
      sc->kappa_lo = (float)(key_th - temp * (double)F_BAND_EPS);
      sc->clamp_key = (float)(Md - 700.0 * temp);
      sc->band_E = ilogb(theta_est) - 1;
      if (!(sc->kappa_lo > sc->clamp_key)) bail = 1;
      sc->bail = bail;
    }
    __syncthreads();
    if (sc->bail) { if (tid == 0) hand_over(P, slow_ws, row); continue; }
    const float kappa_hi = sc->kappa_hi, kappa_lo = sc->kappa_lo, clamp_key = sc->clamp_key;

    auto a_of = [&](float key) -> double {                  // (double(x)/temp) - (double(max)/temp), :128-130
      double x = (double)fmaxf(key, clamp_key);
      if (!UNIT_TEMP) x = __ddiv_rn(x, temp);
      return x - dm;
    };

    // ------------------------------------------------------------------ P1: the fp64 exp pass
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0, accl = 0.0;
    int cnt_hi = 0;
#define F_P1(KEY, OUT, ID, ACC)                                             \
    do {                                                                      \
      const double e_ = ns_exp64_core(a_of(KEY), tab);                        \
      (ACC) += e_;                                                            \
      if ((KEY) >= kappa_hi) { cnt_hi++; (OUT) = __double2float_rz(e_); }     \
      else {                                                                  \
        (OUT) = 0.0f;                                                         \
        if ((KEY) < kappa_lo) accl += e_;                                     \
        else {                                                                \
          const int s_ = atomicAdd(&sc->band_n, 1);                           \
          if (s_ < F_BAND_CAP) { band[s_].id = (ID); band[s_].kept = 0; band[s_].e = e_; } \
        }                                                                     \
      }                                                                       \
    } while (0)
    for (int c = tid; c < W4; c += FT) {
      const float4 v = w4[c];
      const int b = 4 * c - mis;
      float4 o;
      F_P1(v.x, o.x, b, acc0);
      F_P1(v.y, o.y, b + 1, acc1);
      F_P1(v.z, o.z, b + 2, acc2);
      F_P1(v.w, o.w, b + 3, acc3);
      w4[c] = o;
    }
#undef F_P1
    const double sum_all = f_sum_d((acc0 + acc1) + (acc2 + acc3), sc->red);   // softmax normaliser, :130
    const double sum_lo = f_sum_d(accl, sc->red);
    const u64 n_hi = f_reduce_u((u64)cnt_hi, OpAddU(), sc->red);
    const double inv = __ddiv_rn(1.0, sum_all);
    const int nband = sc->band_n;
    const double band_scale = scalbn(1.0, 52 - sc->band_E);
    // ------------------------------------------------------------------ FIX: exact classification
    int my_band_kept = 0;
    if (nband <= F_BAND_CAP && tid < nband) {
      const double e = band[tid].e;
      const bool k = (e * inv) >= thr;                       // p_i >= 1/range, :69
      band[tid].kept = k ? 1 : 0;
      my_band_kept = k ? 1 : 0;
      if (!k) atomicAdd(&sc->band_cut_int, (u64)__double2ull_rz(e * band_scale));   // exact, order-free
    }
    if (tid == 0) {
      // the provisional split is valid iff exp is monotone and both band edges classify as assumed
      const double e_hi = ns_exp64_core(a_of(kappa_hi), tab);
      const double e_lo = ns_exp64_core(a_of(nextafterf(kappa_lo, -INFINITY)), tab);
      int bail = (nband > F_BAND_CAP) || !((e_hi * inv) >= thr) || ((e_lo * inv) >= thr);
      sc->bail = bail;
    }
    const u64 n_band_kept = f_reduce_u((u64)my_band_kept, OpAddU(), sc->red);   // (syncs inside)
    const u64 cand = n_hi + n_band_kept;
    if (sc->bail || !(cand >= 2 && cand <= (u64)P.topk)) {   // rank form (top-k inside the cutoff set) -> exact kernel
      if (tid == 0) hand_over(P, slow_ws, row);
      continue;
    }
    const double sum_bc = (double)sc->band_cut_int * scalbn(1.0, sc->band_E - 52);
    const double S = (sum_all - sum_lo) - sum_bc;            // sum of the kept e_i
    const double C = __ddiv_rn((double)R, S);                // :146
    const double C_lo = C * (1.0 - 2.220446049250313e-16);
    const double C_hi = C * (1.0 + 1.1920928955078125e-07 + 9.094947017729282e-13);
    // bucket shift: every kept element (certain or band) has e >= e(kappa_lo_pred) > 0
    int SH;
    {
      const float e_min = __double2float_rz(ns_exp64_core(a_of(nextafterf(kappa_lo, -INFINITY)), tab));
      const uint32_t span = F_TOP - __float_as_uint(e_min);
      SH = 0;
      while ((span >> SH) > (uint32_t)(F_NB - 1)) ++SH;
    }
    auto bin_of_e = [&](float e32) -> uint32_t { return (F_TOP - __float_as_uint(e32)) >> SH; };

    // ------------------------------------------------------------------ P2: integer bin widths
    u64 qacc = 0;
#define F_P2(E32, ID)                                                        \
    do {                                                                      \
      if ((E32) > 0.0f) {                                                     \
        const double ed_ = (double)(E32);                                     \
        const uint32_t ql_ = (uint32_t)ns_double_as_u64(__fma_rn(ed_, C_lo, magic)); \
        const uint32_t qh_ = (uint32_t)ns_double_as_u64(__fma_rn(ed_, C_hi, magic)); \
        if (ql_ == qh_) { atomicAdd(&hist[bin_of_e(E32)], ql_); qacc += ql_; } \
        else { const int s_ = atomicAdd(&sc->u_n, 1); if (s_ < F_U_CAP) ulist[s_] = (ID); } \
      }                                                                       \
    } while (0)
    for (int c = tid; c < W4; c += FT) {
      const float4 v = w4[c];
      const int b = 4 * c - mis;
      F_P2(v.x, b); F_P2(v.y, b + 1); F_P2(v.z, b + 2); F_P2(v.w, b + 3);
    }
#undef F_P2
    __syncthreads();
    const int nu = sc->u_n;
    if (nu > F_U_CAP) { if (tid == 0) hand_over(P, slow_ws, row); continue; }
    // exact bin width from the original logit (same formula as the exact kernel)
    auto exact_mass = [&](int id) -> u64 {
      const float key = g[id] + 0.0f;
      return (u64)__double2ll_rn(ns_exp64_core(a_of(key), tab) * C);
    };
    for (int u = tid; u < nu; u += FT) {
      const int id = ulist[u];
      const u64 q = exact_mass(id);
      atomicAdd(&hist[bin_of_e(words[id + mis])], (uint32_t)q);
      qacc += q;
    }
    if (tid < nband && band[tid].kept) {
      const double e = band[tid].e;
      const u64 q = (u64)__double2ll_rn(e * C);
      atomicAdd(&hist[bin_of_e(__double2float_rz(e))], (uint32_t)q);
      qacc += q;
    }
    const u64 Q = f_reduce_u(qacc, OpAddU(), sc->red);
    __syncthreads();

    // ------------------------------------------------------------------ SEL helpers
    // bucket of bucket-prefix search: first bucket whose inclusive prefix exceeds tau
    auto locate = [&](u64 tau) {
      constexpr int BPT = F_NB / FT;
      const int lane = tid & 31, warp = tid >> 5;
      u64 local[BPT];
      u64 tsum = 0;
#pragma unroll
      for (int b = 0; b < BPT; ++b) { local[b] = hist[tid * BPT + b]; tsum += local[b]; }
      u64 inc = tsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      __syncthreads();
      if (lane == 31) sc->red[warp] = inc;
      if (tid == 0) { sc->sel_bin = -1; sc->sel_prefix = 0; }
      __syncthreads();
      u64 woff = 0;
      for (int w = 0; w < warp; ++w) woff += sc->red[w];
      u64 excl = woff + inc - tsum;
#pragma unroll
      for (int b = 0; b < BPT; ++b) {
        if (local[b] != 0 && excl <= tau && tau < excl + local[b]) { sc->sel_bin = tid * BPT + b; sc->sel_prefix = excl; }
        excl += local[b];
      }
      __syncthreads();
    };
    // mass in all buckets before bucket tb
    auto prefix_of = [&](int tb) -> u64 {
      constexpr int BPT = F_NB / FT;
      u64 s = 0;
#pragma unroll
      for (int b = 0; b < BPT; ++b) { const int idx = tid * BPT + b; if (idx < tb) s += hist[idx]; }
      return f_reduce_u(s, OpAddU(), sc->red);
    };
    // gather bucket tb: exact order key (original logit, id) and exact mass of every kept element in it
    auto collect = [&](int tb) -> int {
      if (tid == 0) sc->c_n = 0;
      __syncthreads();
      for (int c = tid; c < W4; c += FT) {
        const float4 v = w4[c];
        const int b = 4 * c - mis;
        const float ev[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (bin_of_e(ev[j]) == (uint32_t)tb) {              // e32 == 0 maps far outside the histogram
            const int id = b + j;
            const int s = atomicAdd(&sc->c_n, 1);
            if (s < F_C_CAP) { clist[s].pack = pack_of(g[id] + 0.0f, id); clist[s].w = exact_mass(id); }
          }
        }
      }
      if (tid < nband && band[tid].kept && bin_of_e(__double2float_rz(band[tid].e)) == (uint32_t)tb) {
        const int id = band[tid].id;
        const int s = atomicAdd(&sc->c_n, 1);
        if (s < F_C_CAP) { clist[s].pack = pack_of(g[id] + 0.0f, id); clist[s].w = (u64)__double2ll_rn(band[tid].e * C); }
      }
      __syncthreads();
      int n = sc->c_n;
      if (n > F_C_CAP) { n = F_C_CAP; if (tid == 0 && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW); }
      return n;
    };
    // exact position inside the gathered bucket: by cumulative target tau, or of a given token id
    auto resolve = [&](int n, u64 prefix, bool by_token, u64 tau, int want_id) -> bool {
      if (tid == 0) sc->res_found = 0;
      __syncthreads();
      for (int c = tid; c < n; c += FT) {
        const u64 pc = clist[c].pack, wc = clist[c].w;
        u64 before = prefix;
        for (int o = 0; o < n; ++o) if (clist[o].pack > pc) before += clist[o].w;
        const bool hit = by_token ? (id_of_pack(pc) == want_id) : (wc != 0 && before <= tau && tau < before + wc);
        if (hit) { sc->res_idx = id_of_pack(pc); sc->res_before = before; sc->res_w = wc; sc->res_found = 1; }
      }
      __syncthreads();
      const bool f = sc->res_found != 0;
      return f;
    };
    auto select_tau = [&](u64 tau, int* idx, u64* before, u64* w) -> bool {
      locate(tau);
      const int tb = sc->sel_bin;
      const u64 pref = sc->sel_prefix;
      if (tb < 0) return false;
      const int n = collect(tb);
      const bool f = resolve(n, pref, false, tau, 0);
      *idx = sc->res_idx; *before = sc->res_before; *w = sc->res_w;
      __syncthreads();
      return f;
    };

    // ------------------------------------------------------------------ overfill (:153-158)
    u64 slack;
    bool truncated = false;
    u64 trunc_pack = 0;
    if (Q > R) {
      int j; u64 bj, wj;
      if (select_tau(R, &j, &bj, &wj)) { truncated = true; trunc_pack = pack_of(g[j] + 0.0f, j); slack = R - bj; }
      else slack = 0;
    } else {
      slack = R - Q;
    }
    const u64 top_mass = (u64)__double2ll_rn(C);             // e of the row maximum is exactly 1
    u64 nb, nt;
    if (MODE == MODE_ENC) {
      const int cursor = P.cursor[row];
      const int mlen = P.msg_len[row];
      const u64 window = ns_read_bits(P.msg + (size_t)row * P.msg_stride, cursor, mlen, P.precision);  // :168-171
      const u64 m_rel = window - lo;
      int token;
      if (m_rel < top_mass + slack) {                        // rank 0 absorbs the slack (:158)
        token = top_id; nb = lo; nt = lo + top_mass + slack;
      } else {
        int s; u64 bs, ws;
        if (!select_tau(m_rel - slack, &s, &bs, &ws)) {
          s = top_id; bs = 0; ws = top_mass;
          if (tid == 0 && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW);
        }
        token = s;                                           // :172
        if (s == top_id) { nb = lo; nt = lo + ws + slack; }
        else { nb = lo + bs + slack; nt = nb + ws; }         // :175-176
      }
      if (tid == 0) finish_encode(P, row, slot, token, nb, nt, cand, Q);
    } else {
      int tok = P.token_in[(size_t)row * P.token_stride + slot];
      if (tok < 0 || tok >= V) tok = top_id;
      // is the observed token in the kept set, and in which bucket?
      const float e32t = words[tok + mis];
      int tb = -1;
      if (e32t > 0.0f) tb = (int)bin_of_e(e32t);
      else {
        for (int k = 0; k < nband; ++k)
          if (band[k].id == tok && band[k].kept) tb = (int)bin_of_e(__double2float_rz(band[k].e));
      }
      bool in_range = tb >= 0;
      u64 bs = 0, ws = top_mass;
      int token = top_id;
      if (in_range) {
        const u64 pref = prefix_of(tb);
        const int n = collect(tb);
        if (resolve(n, pref, true, 0, tok)) { bs = sc->res_before; ws = sc->res_w; token = tok; }
        else in_range = false;
        __syncthreads();
        if (in_range && truncated && !(pack_of(g[tok] + 0.0f, tok) > trunc_pack)) { in_range = false; token = top_id; bs = 0; ws = top_mass; }
      }
      if (token == top_id) { nb = lo; nt = lo + ws + slack; }   // :342 / :347-348
      else { nb = lo + bs + slack; nt = nb + ws; }
      if (tid == 0) finish_decode(P, row, slot, in_range, nb, nt, cand, Q);
    }
  }
}
#undef F_ONLINE
