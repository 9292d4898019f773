// ns_coder.cu -- arithmetic-coder step (A) for B independent streams, sm_100a.
//
// Reference behaviour: code_base/arithmetic.py:114-210 (encode loop body) and
// :255-371 (decode loop body); exact step specification in SURVEY.md section 8a.1.
// One CTA owns one stream: the V-wide fp32 logits row lives in shared memory and is
// never sorted.  The reference's "sort, softmax, cut, round, cumsum, search" becomes
//   P1  fp64 exp of every element, fixed-order sum            (softmax normaliser, :130)
//   P2  kept set: p_i >= 1/range, clamped to [2, topk]        (:140-142)
//   P3  integer bin widths q_i = rint(e_i * range / S_kept)   (:146-149), total mass,
//       mass histogram over 2048 monotone key buckets          (replaces :127 + :150)
//   SEL prefix over buckets -> bucket holding the target -> collect that bucket ->
//       exact rank inside it                                   (:153-155 overfill, :172 search)
//   UPD shared-prefix bits + interval rescale                  (:175-190)
// See DESIGN.md for the data layout, the floating-point contract and the roofline.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ns_block.cuh"

namespace {

// Everything a thread needs to turn a key into its exp / probability / bin width.
struct RowMath {
  const float* keys;
  const double* tab;
  int V;
  float m;            // row maximum (fp32, exact)
  double dm;          // double(m)/temp
  double temp;
  bool unit_temp;
  double inv_sum;     // 1 / sum_i e_i                       (softmax, arithmetic.py:130)
  double thr;         // 1 / range                           (:141)
  bool rank_form;     // kept set = top-k0 by order instead of p >= thr
  u64 bound_pack;     // last kept element in the coder's order
  double C;           // range / sum_kept e_i                (:146)

  // (double(x)/temp) - (double(m)/temp), the reference's operation order (:128-130).
  __device__ __forceinline__ double a_of(float key) const {
    double x = (double)key;
    if (!unit_temp) x = __ddiv_rn(x, temp);
    return x - dm;
  }
  __device__ __forceinline__ double e_of(float key) const { return ns_exp64_neg(a_of(key), tab); }
  __device__ __forceinline__ bool kept(float key, int id, double e) const {
    if (rank_form) return pack_of(key, id) >= bound_pack;
    return (e * inv_sum) >= thr;
  }
  // integer bin width of element i, 0 when not kept (:146-149, round half to even)
  __device__ __forceinline__ u64 mass(float key, int id) const {
    if (rank_form && pack_of(key, id) < bound_pack) return 0;
    double e = e_of(key);
    if (!rank_form && !((e * inv_sum) >= thr)) return 0;
    return (u64)__double2ll_rn(e * C);
  }
};

enum { MODE_ENC = 0, MODE_DEC = 1, MODE_DEBUG = 2 };

// ---- per-stream epilogues (one thread) ---------------------------------------------------------
// encode: consume the shared prefix, rescale, emit the token (code_base/arithmetic.py:179-203)
__device__ __forceinline__ void finish_encode(const ns_ac_params& P, int row, int slot, int token,
                                              u64 nb, u64 nt, u64 k0, u64 Q, int cursor, int mlen) {
  uint64_t nlo, nhi;
  const int n = ns_interval_update(nb, nt, P.precision, &nlo, &nhi);   // :179-190
  P.lo[row] = nlo; P.hi[row] = nhi;
  const int nc = cursor + n;                                 // :184
  P.cursor[row] = nc;
  P.token_out[(size_t)row * P.token_stride + slot] = token;  // :202
  if (P.ntok) P.ntok[row] = slot + 1;
  if (P.nbits_out) P.nbits_out[row] = (uint8_t)n;
  if (P.phase && nc >= mlen) P.phase[row] = P.finish_sent ? NS_PHASE_TAIL : NS_PHASE_DONE;   // :114
  if (P.trace) { uint64_t* t = P.trace + (size_t)row * 4; t[0] = nb; t[1] = nt; t[2] = k0; t[3] = Q; }
}

// finish_sent tail: rank-0 token, interval untouched (arithmetic.py:135-137)
__device__ __forceinline__ void finish_tail(const ns_ac_params& P, int row, int slot, int top_id) {
  P.token_out[(size_t)row * P.token_stride + slot] = top_id;
  if (P.ntok) P.ntok[row] = slot + 1;
  if (P.nbits_out) P.nbits_out[row] = 0;
  if (P.phase && P.sent_end && P.sent_end[top_id]) P.phase[row] = NS_PHASE_DONE;
}

// decode: emit the shared prefix (all `precision` bits of new_bottom on the last token), rescale
// (code_base/arithmetic.py:351-366)
__device__ __forceinline__ void finish_decode(const ns_ac_params& P, int row, int slot, bool in_range,
                                              u64 nb, u64 nt, u64 k0, u64 Q) {
  uint64_t nlo, nhi;
  const int n = ns_interval_update(nb, nt, P.precision, &nlo, &nhi);
  P.lo[row] = nlo; P.hi[row] = nhi;
  const bool last = P.ntok_total ? (slot == P.ntok_total[row] - 1) : (P.is_last && P.is_last[row]);
  if (P.ntok) P.ntok[row] = slot + 1;
  if (P.phase && P.ntok_total && slot + 1 >= P.ntok_total[row]) P.phase[row] = NS_PHASE_DONE;
  const int olen = P.out_len[row];
  uint32_t* ob = P.out_bits + (size_t)row * P.out_stride;
  if (last) {                                                // :356-357
    ns_write_bits(ob, olen, nb, P.precision);
    P.out_len[row] = olen + P.precision;
  } else {                                                   // :359
    if (n > 0) ns_write_bits(ob, olen, (nt - 1) >> (P.precision - n), n);
    P.out_len[row] = olen + n;
  }
  if (P.nbits_out) P.nbits_out[row] = (uint8_t)n;
  if (!in_range && P.status) atomicOr(&P.status[row], NS_ST_OUT_OF_RANGE);
  if (P.trace) { uint64_t* t = P.trace + (size_t)row * 4; t[0] = nb; t[1] = nt; t[2] = k0; t[3] = Q; }
}

template <int MODE, typename HistT>
__device__ void ac_exact_row(const ns_ac_params& P, const int row, u64* dbg_q, u64* dbg_meta, unsigned char* smem_raw) {
  constexpr int NB = HIST_BYTES / (int)sizeof(HistT);
  double* tab = reinterpret_cast<double*>(smem_raw);
  HistT* hist = reinterpret_cast<HistT*>(smem_raw + NS_EXP_N * 8);
  ListEntry* list = reinterpret_cast<ListEntry*>(smem_raw + NS_EXP_N * 8 + HIST_BYTES);
  Scalars* sc = reinterpret_cast<Scalars*>(smem_raw + NS_EXP_N * 8 + HIST_BYTES + LIST_CAP * sizeof(ListEntry));
  float* keys_base = reinterpret_cast<float*>(smem_raw + FIXED_BYTES);

  const int tid = threadIdx.x;
  const int V = P.V;
  uint8_t phase = P.phase ? P.phase[row] : (uint8_t)NS_PHASE_CODING;
  if (phase == NS_PHASE_DONE) return;
  if (MODE != MODE_ENC) phase = NS_PHASE_CODING;
  const int slot = P.ntok ? P.ntok[row] : 0;
  if (MODE == MODE_ENC && P.ntok && slot >= P.token_cap) {
    if (tid == 0) {
      if (P.phase) P.phase[row] = NS_PHASE_DONE;
      if (P.status) atomicOr(&P.status[row], NS_ST_TOKEN_OVERFLOW);
    }
    return;
  }
  if (MODE == MODE_DEC && P.ntok_total && slot >= P.ntok_total[row]) {
    if (tid == 0 && P.phase) P.phase[row] = NS_PHASE_DONE;
    return;
  }

  // ---- stage the row: 128-bit loads, shared copy keeps the global 16-byte phase --------
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  float* keys = keys_base + mis;
  {
    int head = (4 - mis) & 3;
    if (head > V) head = V;
    if (tid < head) keys[tid] = g[tid] + 0.0f;               // +0.0f folds -0 into +0
    const int nvec = (V - head) >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g + head);
    float4* s4 = reinterpret_cast<float4*>(keys + head);
    for (int i = tid; i < nvec; i += NT) {
      float4 v = __ldg(g4 + i);
      v.x += 0.0f; v.y += 0.0f; v.z += 0.0f; v.w += 0.0f;
      s4[i] = v;
    }
    const int done = head + (nvec << 2);
    if (tid < V - done) keys[done + tid] = g[done + tid] + 0.0f;
  }
  for (int i = tid; i < NS_EXP_N; i += NT) tab[i] = c_exp_tab[i];
  for (int i = tid; i < NB; i += NT) hist[i] = 0;
  if (tid == 0) sc->list_count = 0;
  __syncthreads();
  // forbidden tokens (code_base/arithmetic.py:124-125 / :265-266): probability exactly 0
  if (tid < 2) {
    int id = P.mask_id[tid];
    if (id >= 0 && id < V) keys[id] = -INFINITY;
  }
  __syncthreads();

  // ---- row maximum (with lowest id among ties) and lowest finite key ---------------------
  u64 pmax = 0, pmin = ~0ull;
  for (int i = tid; i < V; i += NT) {
    float k = keys[i];
    u64 p = pack_of(k, i);
    pmax = p > pmax ? p : pmax;
    if (k > -INFINITY) pmin = p < pmin ? p : pmin;
  }
  pmax = block_reduce_u(pmax, OpMaxU(), sc->red);
  pmin = block_reduce_u(pmin, OpMinU(), sc->red);
  const float m = key_of_pack(pmax);
  const int top_id = id_of_pack(pmax);

  if (MODE == MODE_ENC && phase == NS_PHASE_TAIL) {
    // finish_sent tail: rank-0 token, interval untouched (arithmetic.py:135-137)
    if (tid == 0) finish_tail(P, row, slot, top_id);
    return;
  }

  const u64 lo = P.lo[row], hi = P.hi[row];
  const u64 R = hi - lo;                                    // arithmetic.py:140

  RowMath rm;
  rm.keys = keys; rm.tab = tab; rm.V = V; rm.m = m; rm.temp = P.temp;
  rm.unit_temp = (P.temp == 1.0);
  rm.dm = rm.unit_temp ? (double)m : __ddiv_rn((double)m, P.temp);
  rm.thr = __ddiv_rn(1.0, (double)R);                       // :141
  rm.rank_form = false; rm.bound_pack = 0; rm.C = 0.0; rm.inv_sum = 0.0;

  // ---- P1: softmax normaliser ------------------------------------------------------------
  double sum_e = 0.0;
  {
    double acc0 = 0.0, acc1 = 0.0;
    int i = tid;
    for (; i + NT < V; i += 2 * NT) {
      acc0 += rm.e_of(keys[i]);
      acc1 += rm.e_of(keys[i + NT]);
    }
    if (i < V) acc0 += rm.e_of(keys[i]);
    const double sum = block_sum_d(acc0 + acc1, sc->red);
    rm.inv_sum = __ddiv_rn(1.0, sum);
    sum_e = sum;
  }

  // ---- P2: kept set -------------------------------------------------------------------------
  u64 cand;
  double S;
  u64 bound = ~0ull;
  {
    u64 cnt = 0;
    double acc = 0.0;
    for (int i = tid; i < V; i += NT) {
      const float k = keys[i];
      const double e = rm.e_of(k);
      if ((e * rm.inv_sum) >= rm.thr) {
        cnt += 1;
        acc += e;
        u64 p = pack_of(k, i);
        bound = p < bound ? p : bound;
      }
    }
    cand = block_reduce_u(cnt, OpAddU(), sc->red);
    bound = block_reduce_u(bound, OpMinU(), sc->red);
    S = block_sum_d(acc, sc->red);
  }
  u64 k0 = cand < 2 ? 2 : cand;                              // :75  min(max(2, cand), topk)
  if (k0 > (u64)P.topk) k0 = (u64)P.topk;
  float span_key = key_of_pack(bound);
  if (!(cand >= 2 && cand <= (u64)P.topk)) {
    // rank form: the k0-th element of the order bounds the kept set.  Count histogram over
    // all finite keys, then the same locate/collect/resolve as for masses.
    rm.rank_form = true;
    const float lowest = key_of_pack(pmin);
    const float span = m - lowest;
    const float scale = span > 0.0f ? (float)NB / span : 0.0f;
    for (int i = tid; i < V; i += NT) {
      const float k = keys[i];
      if (k > -INFINITY) atomicAdd(&hist[bin_of(k, m, scale, NB)], (HistT)1);
    }
    __syncthreads();
    sel_locate<HistT, NB>(hist, k0 - 1, sc);
    const int tb = sc->sel_bin;
    const u64 tprefix = sc->sel_prefix;
    if (tb >= 0) {
      for (int i = tid; i < V; i += NT) {
        const float k = keys[i];
        if (k > -INFINITY && bin_of(k, m, scale, NB) == tb) {
          int slot = atomicAdd(&sc->list_count, 1);
          if (slot < LIST_CAP) { list[slot].pack = pack_of(k, i); list[slot].w = 1; }
        }
      }
    }
    __syncthreads();
    int n = sc->list_count;
    if (n > LIST_CAP) { n = LIST_CAP; if (tid == 0 && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW); }
    sel_resolve(list, n, k0 - 1, tprefix, sc);
    // fewer finite keys than k0 (tiny vocabularies): keep them all
    rm.bound_pack = sc->res_found ? pack_of(keys[sc->res_idx], sc->res_idx) : pmin;
    __syncthreads();
    for (int i = tid; i < NB; i += NT) hist[i] = 0;
    if (tid == 0) sc->list_count = 0;
    double acc = 0.0;
    for (int i = tid; i < V; i += NT) {
      const float k = keys[i];
      if (pack_of(k, i) >= rm.bound_pack) acc += rm.e_of(k);
    }
    S = block_sum_d(acc, sc->red);
    span_key = key_of_pack(rm.bound_pack);
  } else {
    rm.bound_pack = bound;
  }
  rm.C = __ddiv_rn((double)R, S);                            // :146  p/sum(p)*range

  // ---- P3: bin widths, total mass, mass histogram; decode: mass before the observed token --
  const float mspan = m - span_key;
  const float mscale = mspan > 0.0f ? (float)NB / mspan : 0.0f;
  int tok = -1;
  u64 tok_pack = 0;
  if (MODE == MODE_DEC) {
    tok = P.token_in[(size_t)row * P.token_stride + slot];
    if (tok < 0 || tok >= V) tok = top_id;
    tok_pack = pack_of(keys[tok], tok);
  }
  u64 Q;
  u64 mass_before_tok = 0;
  {
    u64 q_acc = 0, b_acc = 0;
    for (int i = tid; i < V; i += NT) {
      const float k = keys[i];
      const u64 w = rm.mass(k, i);
      if (MODE == MODE_DEBUG) dbg_q[(size_t)row * V + i] = w;
      if (w != 0) {
        atomicAdd(&hist[bin_of(k, m, mscale, NB)], (HistT)w);
        q_acc += w;
        if (MODE == MODE_DEC && pack_of(k, i) > tok_pack) b_acc += w;
      }
    }
    Q = block_reduce_u(q_acc, OpAddU(), sc->red);
    if (MODE == MODE_DEC) mass_before_tok = block_reduce_u(b_acc, OpAddU(), sc->red);
  }
  __syncthreads();

  // selection helper (uniform control flow across the CTA)
  auto select = [&](u64 tau, int* idx, u64* before, u64* w) -> bool {
    sel_locate<HistT, NB>(hist, tau, sc);
    const int tb = sc->sel_bin;
    const u64 tprefix = sc->sel_prefix;
    if (tid == 0) sc->list_count = 0;
    __syncthreads();
    if (tb < 0) return false;
    for (int i = tid; i < V; i += NT) {
      const float k = keys[i];
      if (bin_of(k, m, mscale, NB) == tb) {
        const u64 wi = rm.mass(k, i);
        if (wi != 0) {
          int slot = atomicAdd(&sc->list_count, 1);
          if (slot < LIST_CAP) { list[slot].pack = pack_of(k, i); list[slot].w = wi; }
        }
      }
    }
    __syncthreads();
    int n = sc->list_count;
    if (n > LIST_CAP) { n = LIST_CAP; if (tid == 0 && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW); }
    sel_resolve(list, n, tau, tprefix, sc);
    const bool found = sc->res_found != 0;
    *idx = sc->res_idx; *before = sc->res_before; *w = sc->res_w;
    __syncthreads();
    return found;
  };

  // ---- overfill: drop the tail of the order once the running total exceeds the range ------
  u64 slack;
  bool truncated = false;
  u64 trunc_pack = 0;            // first dropped element
  if (Q > R) {                                               // :153-155
    int j; u64 bj, wj;
    if (select(R, &j, &bj, &wj)) {
      truncated = true;
      trunc_pack = pack_of(keys[j], j);
      slack = R - bj;                                        // :158
    } else {
      slack = 0;
    }
  } else {
    slack = R - Q;                                           // :158
  }

  if (MODE == MODE_DEBUG) {
    if (truncated) {
      for (int i = tid; i < V; i += NT)
        if (pack_of(keys[i], i) <= trunc_pack) dbg_q[(size_t)row * V + i] = 0;
    }
    if (tid == 0) {
      dbg_meta[row * 4 + 0] = k0;
      dbg_meta[row * 4 + 1] = slack;
      dbg_meta[row * 4 + 2] = Q;
      dbg_meta[row * 4 + 3] = R;
    }
    return;
  }

  const u64 top_mass = rm.mass(keys[top_id], top_id);
  u64 nb, nt;
  int token;
  if (MODE == MODE_ENC) {
    const int cursor = P.cursor[row];
    const int mlen = P.msg_len[row];
    const u64 window = ns_read_bits(P.msg + (size_t)row * P.msg_stride, cursor, mlen, P.precision);  // :168-171
    const u64 m_rel = window - lo;
    if (m_rel < top_mass + slack) {                          // rank 0 absorbs the slack (:158)
      token = top_id; nb = lo; nt = lo + top_mass + slack;
    } else {
      int s; u64 bs, ws;
      if (!select(m_rel - slack, &s, &bs, &ws)) { s = top_id; bs = 0; ws = top_mass; if (tid == 0 && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW); }
      token = s;                                             // :172
      if (s == top_id) { nb = lo; nt = lo + ws + slack; }
      else { nb = lo + bs + slack; nt = nb + ws; }           // :175-176
    }
    if (P.stats) {
      // per-step statistics of the reference (arithmetic.py:131-132,192-198; utils.py:32-40):
      // log p(selected) under the untempered softmax, KL(q_hat || p) in bits over the kept bins,
      // entropy of the tempered distribution in bits.
      __syncthreads();
      const double md = (double)m;
      double s1 = sum_e;
      if (!rm.unit_temp) {
        double acc = 0.0;
        for (int i = tid; i < V; i += NT) acc += ns_exp64_neg((double)keys[i] - md, tab);
        s1 = block_sum_d(acc, sc->red);
      }
      const double ln_s1 = log(s1), ln_se = log(sum_e);
      double kl = 0.0, ent = 0.0;
      for (int i = tid; i < V; i += NT) {
        const float k = keys[i];
        const double a = rm.a_of(k);
        const double e = ns_exp64_neg(a, tab);
        const double pt = e * rm.inv_sum;
        if (pt != 0.0) ent += pt * (a - ln_se) / 0.69315;                  // utils.py:38
        u64 w = rm.mass(k, i);
        if (truncated && pack_of(k, i) <= trunc_pack) w = 0;
        if (i == top_id) w += slack;                                       // :158
        if (w != 0) {
          const double qh = (double)w / (double)R;                         // :195
          kl += qh * (log(qh) - (((double)k - md) - ln_s1)) / 0.69315;     // utils.py:33
        }
      }
      kl = block_sum_d(kl, sc->red);
      ent = block_sum_d(ent, sc->red);
      if (tid == 0) {
        double* st = P.stats + (size_t)row * 3;
        st[0] = ((double)keys[token] - md) - ln_s1;                        // log_probs[selection], :193
        st[1] = kl;
        st[2] = -ent;
      }
    }
    if (tid == 0) finish_encode(P, row, slot, token, nb, nt, k0, Q, cursor, mlen);
  } else {
    // decode: rank of the observed token = mass in front of it (:298)
    const u64 wt = rm.mass(keys[tok], tok);
    bool in_range = (wt != 0 || rm.kept(keys[tok], tok, rm.e_of(keys[tok]))) && (!truncated || tok_pack > trunc_pack);
    token = tok;
    u64 bs = mass_before_tok, ws = wt;
    if (!in_range) { token = top_id; bs = 0; ws = top_mass; }                 // :342 rank = 0
    if (token == top_id) { nb = lo; nt = lo + ws + slack; }
    else { nb = lo + bs + slack; nt = nb + ws; }                            // :347-348
    if (tid == 0) finish_decode(P, row, slot, in_range, nb, nt, k0, Q);
  }
}

// One CTA per row (slow_ws == nullptr), or a small persistent grid draining the rows the fast
// kernel queued in slow_ws = {count, done, rows...}; the last CTA to finish resets the queue.
template <int MODE, typename HistT>
__global__ void __launch_bounds__(NT, 1) ac_step_kernel(ns_ac_params P, u64* dbg_q, u64* dbg_meta, int32_t* slow_ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (!slow_ws) {
    ac_exact_row<MODE, HistT>(P, blockIdx.x, dbg_q, dbg_meta, smem_raw);
    return;
  }
  const int count = slow_ws[0];
  for (int it = blockIdx.x; it < count; it += gridDim.x) {
    __syncthreads();
    ac_exact_row<MODE, HistT>(P, slow_ws[2 + it], dbg_q, dbg_meta, smem_raw);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int d = atomicAdd(&slow_ws[1], 1);
    if (d == (int)gridDim.x - 1) { slow_ws[0] = 0; slow_ws[1] = 0; __threadfence(); }
  }
}

// the single-row throughput kernel (carries the rank form of the cutoff; see the head of ns_fast.cuh)
namespace nsf_smem {
#define NSF_FT 512
#define NSF_MIN_CTAS 1
#include "ns_fast.cuh"
#undef NSF_FT
#undef NSF_MIN_CTAS
}  // namespace nsf_smem
// rank form of the cutoff without a resident row: one 512-thread CTA per row, two per SM
namespace nst {
#include "ns_topk.cuh"
}  // namespace nst
// lean threshold-form kernel: one 1024-thread CTA per SM, one tight loop per sweep
namespace nsl {
#include "ns_lean.cuh"
}  // namespace nsl

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
thread_local char g_err[256] = "";

int set_err(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int validate(const ns_ac_params* p, int mode) {
  if (!p) return set_err(NS_E_NULL, "params is NULL");
  if (!p->logits || !p->lo || !p->hi) return set_err(NS_E_NULL, "logits/lo/hi is NULL");
  if (p->B < 0 || p->V < 4) return set_err(NS_E_RANGE, "B < 0 or V < 4");
  if (p->V > MAX_VOCAB) return set_err(NS_E_VOCAB, "V exceeds the shared-memory row capacity (ns_ac_max_vocab)");
  if (p->ld < p->V) return set_err(NS_E_RANGE, "ld < V");
  if (p->precision < 2 || p->precision > 48) return set_err(NS_E_RANGE, "precision must be in [2, 48]");
  if (p->topk < 1) return set_err(NS_E_RANGE, "topk must be >= 1");
  if (!(p->temp > 0.0)) return set_err(NS_E_RANGE, "temp must be > 0");
  if (((uintptr_t)p->logits & 3u) != 0) return set_err(NS_E_ALIGN, "logits not 4-byte aligned");
  if (mode == MODE_ENC) {
    if (!p->msg || !p->msg_len || !p->cursor || !p->token_out) return set_err(NS_E_NULL, "encode needs msg/msg_len/cursor/token_out");
  } else if (mode == MODE_DEC) {
    if (!p->token_in || !p->out_bits || !p->out_len) return set_err(NS_E_NULL, "decode needs token_in/out_bits/out_len");
  }
  return NS_OK;
}

// SM count and the shared-memory opt-in are properties of the CURRENT device: asked per call (both are cheap host-side
// lookups), so a process that drives several GPUs or several host threads gets the right answer on each.
int num_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
    return n;
  return 148;
}

template <typename K>
int configure(K kernel, bool*) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  if (e != cudaSuccess) { set_err((int)e, cudaGetErrorString(e)); return e == cudaErrorInvalidDeviceFunction ? NS_E_NODEVICE : (int)e; }
  return NS_OK;
}

int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err((int)e, cudaGetErrorString(e));
  return NS_OK;
}

// exact kernel: one CTA per row, or (slow_ws given) a 32-CTA grid draining the hand-over queue
template <int MODE, typename HistT>
int launch_exact(const ns_ac_params* p, u64* dbg_q, u64* dbg_meta, int32_t* slow_ws, cudaStream_t st) {
  const int smem = FIXED_BYTES + (p->V + 8) * 4;
  static bool configured = false;
  int rc = configure(ac_step_kernel<MODE, HistT>, &configured);
  if (rc != NS_OK) return rc;
  if (p->B == 0) return NS_OK;
  const int grid = slow_ws ? (p->B < 32 ? p->B : 32) : p->B;
  ac_step_kernel<MODE, HistT><<<grid, NT, smem, st>>>(*p, dbg_q, dbg_meta, slow_ws);
  return check_launch();
}

// rank form without a resident row: one CTA per row; rows it does not carry land in p->rank_ws
template <bool UNIT, int MODE>
int launch_topk(const ns_ac_params* p, cudaStream_t st) {
  if (p->B == 0) return NS_OK;
  nst::ac_topk_stream_kernel<UNIT, MODE><<<p->B, nst::KT, 0, st>>>(*p, p->rank_ws);
  return check_launch();
}

// fast kernel: persistent, one CTA per SM; `rows` = work list (p->rank_ws) or nullptr for all rows
template <bool UNIT, int MODE, bool RANK>
int launch_fast(const ns_ac_params* p, cudaStream_t st, int32_t* rows = nullptr) {
  const int smem = FIXED_BYTES + (p->V + 8) * 4;
  static bool configured = false;
  int rc = configure(nsf_smem::ac_fast_kernel<UNIT, MODE, RANK>, &configured);
  if (rc != NS_OK) return rc;
  if (p->B == 0) return NS_OK;
  const int sms = num_sms();
  const int grid = p->B < sms ? p->B : sms;
  nsf_smem::ac_fast_kernel<UNIT, MODE, RANK><<<grid, nsf_smem::FT, smem, st>>>(*p, p->slow_ws, rows);
  return check_launch();
}

// lean kernel: persistent, one 1024-thread CTA per SM
template <bool UNIT, int MODE>
int launch_lean(const ns_ac_params* p, cudaStream_t st) {
  const int smem = nsl::L_FIXED + (p->V + 8) * 4;
  static bool configured = false;
  int rc = configure(nsl::ac_lean_kernel<UNIT, MODE, false>, &configured);
  if (rc == NS_OK && p->prof) rc = configure(nsl::ac_lean_kernel<UNIT, MODE, true>, &configured);
  if (rc != NS_OK) return rc;
  if (p->B == 0) return NS_OK;
  const int sms = num_sms();
  const int grid = p->B < sms ? p->B : sms;
  if (p->prof) nsl::ac_lean_kernel<UNIT, MODE, true><<<grid, nsl::LT, smem, st>>>(*p, p->slow_ws);   // with phase timers
  else nsl::ac_lean_kernel<UNIT, MODE, false><<<grid, nsl::LT, smem, st>>>(*p, p->slow_ws);
  return check_launch();
}

template <int MODE>
int dispatch(const ns_ac_params* p, u64* dbg_q, u64* dbg_meta, void* stream) {
  int rc = validate(p, MODE);
  if (rc != NS_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p->precision > 31) return launch_exact<MODE, u64>(p, dbg_q, dbg_meta, nullptr, st);
  if (MODE == MODE_DEBUG || p->slow_ws == nullptr || p->force_exact || p->V < nsf_smem::F_MIN_VOCAB || p->stats != nullptr)
    return launch_exact<MODE, uint32_t>(p, dbg_q, dbg_meta, nullptr, st);
  // throughput path: fast kernel, then the exact kernel on whatever it handed over
  // (top-k small enough to bind and to fit the rank-form lists: the instantiation that carries that path)
  constexpr int M2 = MODE == MODE_DEBUG ? MODE_ENC : MODE;
  const bool rank = p->topk >= 2 && p->topk < p->V && p->topk <= nsf_smem::F_K_CAP;
  const bool lean = !rank && p->variant == 0 && p->V >= nsl::L_MIN_VOCAB && p->V <= nsl::L_MAX_VOCAB;
  // rank form: the sweep kernel first (when the caller gave it a work list), then the row-resident kernel on the rows
  // that one queued (not certainly in rank form, finish_sent tails, degenerate rows)
  const bool topk_sweep = rank && p->variant == 0 && p->rank_ws != nullptr && nst::k_shape_ok(p->V, p->topk);
  if (topk_sweep) {
    rc = (p->temp == 1.0) ? launch_topk<true, M2>(p, st) : launch_topk<false, M2>(p, st);
    if (rc != NS_OK) return rc;
    rc = (p->temp == 1.0) ? launch_fast<true, M2, true>(p, st, p->rank_ws) : launch_fast<false, M2, true>(p, st, p->rank_ws);
  }
  else if (rank) rc = (p->temp == 1.0) ? launch_fast<true, M2, true>(p, st) : launch_fast<false, M2, true>(p, st);
  else if (lean) rc = (p->temp == 1.0) ? launch_lean<true, M2>(p, st) : launch_lean<false, M2>(p, st);
  else rc = (p->temp == 1.0) ? launch_fast<true, M2, false>(p, st) : launch_fast<false, M2, false>(p, st);
  if (rc != NS_OK) return rc;
  return launch_exact<MODE, uint32_t>(p, dbg_q, dbg_meta, p->slow_ws, st);
}

}  // namespace

extern "C" {

int ns_version(void) { return NS_ABI_VERSION; }
const char* ns_last_error_string(void) { return g_err; }
int ns_ac_max_vocab(void) { return MAX_VOCAB; }
int ns_sizeof_ac_params(void) { return (int)sizeof(ns_ac_params); }

int ns_ac_encode_step(const ns_ac_params* p, void* cuda_stream) { return dispatch<MODE_ENC>(p, nullptr, nullptr, cuda_stream); }
int ns_ac_decode_step(const ns_ac_params* p, void* cuda_stream) { return dispatch<MODE_DEC>(p, nullptr, nullptr, cuda_stream); }
int ns_ac_debug_bins(const ns_ac_params* p, uint64_t* q_out, uint64_t* meta_out, void* cuda_stream) {
  if (!q_out || !meta_out) return set_err(NS_E_NULL, "q_out/meta_out is NULL");
  return dispatch<MODE_DEBUG>(p, (u64*)q_out, (u64*)meta_out, cuda_stream);
}

}  // extern "C"
