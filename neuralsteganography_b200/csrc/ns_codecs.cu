// ns_codecs.cu -- the reference's comparison codecs, one CTA per stream, sm_100a.
//   rank    src/neuralstego/codec/arithmetic.py:122-231, :370-385   (what load_lm("gpt2-fa") runs)
//   huffman code_base/huffman_baseline.py:7-71, :73-165 ; code_base/huffman.py:12-76
//   bins    code_base/block_baseline.py:26-97, :99-189
// The row is staged in shared memory once (rank, Huffman) or streamed (bins).  Order everywhere:
// larger logit first, equal logits by lower token id.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ns_block.cuh"

namespace {

constexpr int CODEC_LIST_CAP = LIST_CAP;     // 512 gathered elements (Huffman: 2^bits_per_word <= 512)
constexpr int HUF_MAX_LEAVES = 512;

enum { K_RANK_ENC = 0, K_RANK_DEC, K_HUF_ENC, K_HUF_DEC, K_BINS_ENC, K_BINS_DEC };

struct HufNode { float freq; int left, right, parent; };   // leaves: index < n, left = right = -1

struct CodecShared {
  unsigned char* raw;
  uint32_t* hist;
  ListEntry* list;
  Scalars* sc;
  float* keys;
};

__device__ __forceinline__ CodecShared carve(unsigned char* smem_raw) {
  CodecShared s;
  s.raw = smem_raw;
  s.hist = reinterpret_cast<uint32_t*>(smem_raw + NS_EXP_N * 8);
  s.list = reinterpret_cast<ListEntry*>(smem_raw + NS_EXP_N * 8 + HIST_BYTES);
  s.sc = reinterpret_cast<Scalars*>(smem_raw + NS_EXP_N * 8 + HIST_BYTES + LIST_CAP * sizeof(ListEntry));
  s.keys = reinterpret_cast<float*>(smem_raw + FIXED_BYTES);
  return s;
}

// ------------------------------------------------------------------------------------------
// bulk row staging: the copy engine writes the row into shared memory in C_PIECES pieces while the
// CTA reduces its extent piece by piece (one pass instead of load + store + extent pass)
// ------------------------------------------------------------------------------------------
constexpr int C_PIECES = 2;
__device__ __forceinline__ uint32_t c_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void c_mbar_init(u64* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(c_saddr(bar)), "r"(count));
}
__device__ __forceinline__ void c_mbar_expect_tx(u64* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(c_saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void c_bulk_g2s(void* dst, const void* src, uint32_t bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(c_saddr(dst)), "l"(src), "r"(bytes), "r"(c_saddr(bar)) : "memory");
}
__device__ __forceinline__ void c_mbar_wait(u64* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" :: "r"(c_saddr(bar)), "r"(parity) : "memory");
}

// Stage row `row` (same 16-byte phase as in global memory; -0 folded into +0; forbidden tokens at -1e10 when
// use_mask) and return its extent: *pmax = pack of the largest key (lowest id among equals), *pmin = pack of
// the smallest key above -1e9 (forbidden tokens and -inf do not stretch the bucket range).  One CTA, one row.
__device__ float* stage_row_bulk(const ns_codec_params& P, int row, CodecShared& sm, bool use_mask, u64* pmax, u64* pmin) {
  constexpr int NB = HIST_BYTES / 4;
  const int tid = threadIdx.x, V = P.V;
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  const int mis = (int)(((uintptr_t)g & 15u) >> 2);
  float* keys = sm.keys + mis;                               // element id lives at keys[id]
  float4* k4 = reinterpret_cast<float4*>(sm.keys);
  const int W4 = (mis + V + 3) >> 2;
  const int NI = W4 - 2;                                     // interior chunks 1 .. W4-2: wholly inside the row
  const int PC = ((NI > 0 ? NI : 0) + C_PIECES - 1) / C_PIECES;
  u64* bar = sm.sc->bar;
  if (tid == 0) {
    for (int k = 0; k < C_PIECES; ++k) c_mbar_init(&bar[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const char* src = reinterpret_cast<const char*>(g - mis) + 16;
    char* dst = reinterpret_cast<char*>(k4 + 1);
    for (int k = 0; k < C_PIECES; ++k) {
      const int c0 = k * PC;
      int n = NI - c0;
      if (n > PC) n = PC;
      if (n > 0) {
        c_mbar_expect_tx(&bar[k], (uint32_t)n * 16u);
        c_bulk_g2s(dst + (size_t)c0 * 16, src + (size_t)c0 * 16, (uint32_t)n * 16u, &bar[k]);
      }
    }
    sm.sc->list_count = 0;
  }
  for (int i = tid; i < NB; i += NT) sm.hist[i] = 0;
  const int mk0 = (use_mask && P.mask_id[0] >= 0 && P.mask_id[0] < V) ? P.mask_id[0] : -8;
  const int mk1 = (use_mask && P.mask_id[1] >= 0 && P.mask_id[1] < V) ? P.mask_id[1] : -8;
  // the two edge chunks may straddle the row ends: element-wise, by the eight lanes that also reduce them
  float bk = -INFINITY, lk = INFINITY;                       // this thread's max key (first id among equals), min key
  int bi = 0x7fffffff;
  auto take = [&](float k, int id) {
    if (k > bk || bi == 0x7fffffff) { bk = k; bi = id; }     // ids ascend within a thread
    if (k > -1e9f) lk = fminf(lk, k);
  };
  float ek = 0.0f;
  int eid = -1;
  if (tid < 8) {
    const int c = tid < 4 ? 0 : W4 - 1;
    const int id = 4 * c - mis + (tid & 3);
    if (id >= 0 && id < V) {
      float k = g[id] + 0.0f;
      if (id == mk0 || id == mk1) k = -1e10f;                // huffman_baseline.py:26-27
      keys[id] = k;
      ek = k; eid = id;
    }
  }
  if (tid < 4 && eid >= 0) take(ek, eid);                    // first chunk: the lowest ids
  __syncthreads();                                           // mbarrier init visible before anyone waits
  for (int k = 0; k < C_PIECES; ++k) {
    const int c0 = 1 + k * PC;
    int c1 = c0 + PC;
    if (c1 > 1 + NI) c1 = 1 + NI;
    if (c0 >= c1) break;
    c_mbar_wait(&bar[k], 0);
    for (int c = c0 + tid; c < c1; c += NT) {
      float4 v = k4[c];
      const int id = 4 * c - mis;
      const bool fix = (__float_as_uint(v.x) == 0x80000000u) | (__float_as_uint(v.y) == 0x80000000u) |
                       (__float_as_uint(v.z) == 0x80000000u) | (__float_as_uint(v.w) == 0x80000000u) |
                       (c == ((mk0 + mis) >> 2)) | (c == ((mk1 + mis) >> 2));
      if (fix) {                                             // rare: -0 -> +0, forbidden tokens
        v.x += 0.0f; v.y += 0.0f; v.z += 0.0f; v.w += 0.0f;
        float* f = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) if (id + j == mk0 || id + j == mk1) f[j] = -1e10f;
        k4[c] = v;
      }
      take(v.x, id); take(v.y, id + 1); take(v.z, id + 2); take(v.w, id + 3);
    }
  }
  if (tid >= 4 && eid >= 0) take(ek, eid);                   // last chunk: the highest ids
  u64 a = (bi != 0x7fffffff) ? pack_of(bk, bi) : 0ull;
  u64 b = pack_of(lk, 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const u64 x = __shfl_xor_sync(0xffffffffu, a, o), y = __shfl_xor_sync(0xffffffffu, b, o);
    a = x > a ? x : a;
    b = y < b ? y : b;
  }
  __syncthreads();                                           // fixes of other threads visible; red free
  if ((tid & 31) == 0) { sm.sc->red[tid >> 5] = a; reinterpret_cast<u64*>(sm.list)[tid >> 5] = b; }
  __syncthreads();
  u64 ra = sm.sc->red[tid & 31], rb = reinterpret_cast<u64*>(sm.list)[tid & 31];   // one partial per lane
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const u64 x = __shfl_xor_sync(0xffffffffu, ra, o), y = __shfl_xor_sync(0xffffffffu, rb, o);
    ra = x > ra ? x : ra;
    rb = y < rb ? y : rb;
  }
  *pmax = ra; *pmin = rb;
  __syncthreads();                                           // list scratch free again
  return keys;
}

// The element whose 0-based position in the coder's order is `pos`, from a count histogram over 2048 monotone key
// buckets that the caller has filled (list_count == 0) + exact resolution inside the bucket.  Returns -1 if pos is out of range.
__device__ int element_from_hist(const float* keys, int V, u64 pmax, u64 pmin, u64 pos, CodecShared& sm,
                                 int32_t* status) {
  constexpr int NB = HIST_BYTES / 4;
  const int tid = threadIdx.x;
  const float m = key_of_pack(pmax);
  const float span = m - key_of_pack(pmin);
  const float scale = span > 0.0f ? (float)NB / span : 0.0f;
  __syncthreads();
  sel_locate<uint32_t, NB>(sm.hist, pos, sm.sc);
  const int tb = sm.sc->sel_bin;
  const u64 prefix = sm.sc->sel_prefix;
  if (tb < 0) return -1;
  for (int i = tid; i < V; i += NT) {
    const float k = keys[i];
    if (bin_of(k, m, scale, NB) == tb) {
      const int slot = atomicAdd(&sm.sc->list_count, 1);
      if (slot < CODEC_LIST_CAP) { sm.list[slot].pack = pack_of(k, i); sm.list[slot].w = 1; }
    }
  }
  __syncthreads();
  int n = sm.sc->list_count;
  if (n > CODEC_LIST_CAP) { n = CODEC_LIST_CAP; if (tid == 0 && status) atomicOr(status, NS_ST_BIN_OVERFLOW); }
  sel_resolve(sm.list, n, pos, prefix, sm.sc);
  const int found = sm.sc->res_found ? sm.sc->res_idx : -1;
  __syncthreads();
  return found;
}

__device__ __forceinline__ bool stream_live(const ns_codec_params& P, int row, bool decode, int* slot_out) {
  const uint8_t phase = P.phase ? P.phase[row] : (uint8_t)NS_PHASE_CODING;
  const int slot = P.ntok ? P.ntok[row] : 0;
  *slot_out = slot;
  if (phase == NS_PHASE_DONE) return false;
  if (!decode && P.ntok && slot >= P.token_cap) {
    if (threadIdx.x == 0) {
      if (P.phase) P.phase[row] = NS_PHASE_DONE;
      if (P.status) atomicOr(&P.status[row], NS_ST_TOKEN_OVERFLOW);
    }
    return false;
  }
  if (decode && P.ntok_total && slot >= P.ntok_total[row]) {
    if (threadIdx.x == 0 && P.phase) P.phase[row] = NS_PHASE_DONE;
    return false;
  }
  return true;
}

__device__ __forceinline__ void emit_token(const ns_codec_params& P, int row, int slot, int token, int consumed) {
  const int nc = P.cursor[row] + consumed;
  P.cursor[row] = nc;
  P.token_out[(size_t)row * P.token_stride + slot] = token;
  if (P.ntok) P.ntok[row] = slot + 1;
  if (P.nbits_out) P.nbits_out[row] = (uint8_t)consumed;
  if (P.phase && nc >= P.msg_len[row]) P.phase[row] = NS_PHASE_DONE;
}

__device__ __forceinline__ void emit_bits(const ns_codec_params& P, int row, int slot, u64 value, int count) {
  const int olen = P.out_len[row];
  if (count > 0) ns_write_bits(P.out_bits + (size_t)row * P.out_stride, olen, value, count);
  P.out_len[row] = olen + count;
  if (P.ntok) P.ntok[row] = slot + 1;
  if (P.nbits_out) P.nbits_out[row] = (uint8_t)count;
  if (P.phase && P.ntok_total && slot + 1 >= P.ntok_total[row]) P.phase[row] = NS_PHASE_DONE;
}

// ------------------------------------------------------------------------------------------
// (B) rank codec
// ------------------------------------------------------------------------------------------
template <bool DECODE>
__global__ void __launch_bounds__(NT, 1) rank_kernel(ns_codec_params P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CodecShared sm = carve(smem_raw);
  const int row = blockIdx.x, tid = threadIdx.x, V = P.V;
  int slot;
  if (!stream_live(P, row, DECODE, &slot)) return;
  u64 pmax, pmin;
  const float* keys = stage_row_bulk(P, row, sm, false, &pmax, &pmin);
  const float m = key_of_pack(pmax);
  // tokens with p > 0 (codec/arithmetic.py:372): fp64 softmax underflows below exp(-745)
  const double dm = (double)m / P.temp;
  u64 n_pos = ~0ull;                                                    // quality limits first, the count below
  if (P.topk > 0) n_pos = (u64)P.topk;                                  // quality.py:76-81
  // top_p / min_prob (quality.py:85-96): both keep a prefix of the order, judged on the unfiltered fp64
  // softmax; the renormalisation (:101-103) changes neither the order nor the support
  const bool use_p = P.top_p > 0.0 && P.top_p < 1.0, use_min = P.min_prob > 0.0;
  if (use_p || use_min) {
    double* tab = reinterpret_cast<double*>(sm.raw);
    for (int i = tid; i < NS_EXP_N; i += NT) tab[i] = c_exp_tab[i];
    __syncthreads();
    const double temp = P.temp;
    auto e_of = [&](float k) -> double { return ns_exp64_neg((double)k / temp - dm, tab); };
    double acc = 0.0;
    for (int i = tid; i < V; i += NT) acc += e_of(keys[i]);
    const double inv = 1.0 / block_sum_d(acc, sm.sc->red);              // softmax normaliser, lm/arithmetic.py:73
    if (use_min) {
      u64 c = 0;
      for (int i = tid; i < V; i += NT) c += (e_of(keys[i]) * inv >= P.min_prob) ? 1 : 0;
      const u64 n_min = block_reduce_u(c, OpAddU(), sm.sc->red);
      if (n_min < n_pos) n_pos = n_min;                                 // 0 -> QualityConfigError :98-99, flagged below
    }
    if (use_p) {
      // cutoff = #(positions whose running sum stays below top_p) (:89-90): masses as 2^-62 fixed point (order-free
      // sums), a mass histogram over 1024 monotone key buckets, exact resolution inside the bucket
      constexpr int NB64 = HIST_BYTES / 8;
      u64* h64 = reinterpret_cast<u64*>(sm.hist);
      const double two62 = 4611686018427387904.0;
      const double f_top = ceil(P.top_p * two62);
      const u64 tau = f_top >= 1.0 ? (u64)f_top - 1ull : 0ull;          // first position with sum >= top_p
      const float span = m - key_of_pack(pmin);
      const float scale = span > 0.0f ? (float)NB64 / span : 0.0f;
      __syncthreads();
      for (int i = tid; i < NB64; i += NT) h64[i] = 0ull;
      if (tid == 0) sm.sc->list_count = 0;
      __syncthreads();
      for (int i = tid; i < V; i += NT) {
        const float k = keys[i];
        const u64 f = __double2ull_rz(e_of(k) * inv * two62);
        if (f) atomicAdd(&h64[bin_of(k, m, scale, NB64)], f);
      }
      __syncthreads();
      sel_locate<u64, NB64>(h64, tau, sm.sc);
      const int tb = sm.sc->sel_bin;
      const u64 prefix = sm.sc->sel_prefix;
      if (tb >= 0) {                                                    // else the sum never reaches top_p: keep all
        for (int i = tid; i < V; i += NT) {
          const float k = keys[i];
          if (bin_of(k, m, scale, NB64) == tb) {
            const u64 f = __double2ull_rz(e_of(k) * inv * two62);
            const int s = atomicAdd(&sm.sc->list_count, 1);
            if (s < CODEC_LIST_CAP) { sm.list[s].pack = pack_of(k, i); sm.list[s].w = f; }
          }
        }
        __syncthreads();
        int n = sm.sc->list_count;
        if (n > CODEC_LIST_CAP) { n = CODEC_LIST_CAP; if (tid == 0 && P.status) atomicOr(&P.status[row], NS_ST_BIN_OVERFLOW); }
        sel_resolve(sm.list, n, tau, prefix, sm.sc);
        if (sm.sc->res_found) {
          const u64 tp = pack_of(keys[sm.sc->res_idx], sm.sc->res_idx);
          u64 before = 0;
          for (int i = tid; i < V; i += NT) before += pack_of(keys[i], i) > tp ? 1 : 0;
          const u64 n_p = block_reduce_u(before, OpAddU(), sm.sc->red) + 1ull;   // order[:cutoff + 1], :91
          if (n_p < n_pos) n_pos = n_p;
        }
      }
      __syncthreads();
    }
  }
  // one fused pass: count the tokens with p > 0 and (encode) fill the count histogram of the selection,
  // or (decode) count the tokens ranked before the observed one.  The underflow test is a float compare
  // except within a hair of the boundary, where the fp64 expression decides.
  const float kc = (float)((dm - 745.0) * P.temp);
  const float margin = fmaxf(1e-3f, fabsf(kc) * 1e-5f);
  const float k_yes = kc + margin, k_no = kc - margin;
  auto positive = [&](float k) -> bool {
    if (k > k_yes) return true;
    if (k < k_no) return false;
    return ((double)k / P.temp - dm) >= -745.0;
  };
  constexpr int NB = HIST_BYTES / 4;
  const float span = m - key_of_pack(pmin);
  const float scale = span > 0.0f ? (float)NB / span : 0.0f;
  int tok = 0;
  u64 tp = 0;
  if (DECODE) {
    tok = P.token_in[(size_t)row * P.token_stride + slot];
    if (tok < 0 || tok >= V) tok = id_of_pack(pmax);
    tp = pack_of(keys[tok], tok);
  } else {
    __syncthreads();
    for (int i = tid; i < NB; i += NT) sm.hist[i] = 0;                  // the filters may have used it
    if (tid == 0) sm.sc->list_count = 0;
    __syncthreads();
  }
  // branch-free: certain tokens and "within a hair of the bound" tokens are counted apart; the fp64 test runs
  // only if the second count is not zero (practically never)
  uint32_t cnt = 0, maybe = 0, before = 0;
  for (int i = tid; i < V; i += NT) {
    const float k = keys[i];
    cnt += k > k_yes ? 1u : 0u;
    maybe += (k >= k_no && !(k > k_yes)) ? 1u : 0u;
    if (DECODE) before += pack_of(k, i) > tp ? 1u : 0u;
    else atomicAdd(&sm.hist[bin_of(k, m, scale, NB)], 1u);
  }
  const u64 both = block_reduce_u(((u64)before << 40) | ((u64)maybe << 20) | (u64)cnt, OpAddU(), sm.sc->red);
  u64 total_pos = both & 0xfffffull;
  if ((both >> 20) & 0xfffffull) {
    u64 c2 = 0;
    for (int i = tid; i < V; i += NT) {
      const float k = keys[i];
      if (k >= k_no && !(k > k_yes) && positive(k)) c2 += 1;
    }
    total_pos += block_reduce_u(c2, OpAddU(), sm.sc->red);
  }
  if (total_pos < n_pos) n_pos = total_pos;
  const int capacity = n_pos ? 63 - __clzll((long long)n_pos) : 0;      // floor(log2(n_pos)), :379
  if (capacity <= 0) {                                                  // ArithmeticRangeError :149
    if (tid == 0 && P.status) atomicOr(&P.status[row], NS_ST_OUT_OF_RANGE);
    if (tid == 0 && P.phase) P.phase[row] = NS_PHASE_DONE;
    return;
  }
  if (!DECODE) {
    const int cursor = P.cursor[row], mlen = P.msg_len[row];
    const u64 index = ns_read_bits(P.msg + (size_t)row * P.msg_stride, cursor, mlen, capacity);   // :153-157
    int token = element_from_hist(keys, V, pmax, pmin, index, sm, P.status ? &P.status[row] : nullptr);
    if (token < 0) token = id_of_pack(pmax);
    int consumed = mlen - cursor;
    if (consumed > capacity) consumed = capacity;                       // :155
    if (tid == 0) emit_token(P, row, slot, token, consumed);
  } else {
    const u64 rank = both >> 40;                                        // ranked_tokens.index(token), :211
    if (tid == 0) {
      if (rank >= (1ull << capacity) && P.status) atomicOr(&P.status[row], NS_ST_OUT_OF_RANGE);   // :212-213
      const int total = P.total_bits ? P.total_bits[row] : 0x7fffffff;
      int take = total - P.out_len[row];
      if (take > capacity) take = capacity;
      if (take < 0) take = 0;
      emit_bits(P, row, slot, (rank & ((1ull << capacity) - 1ull)) >> (capacity - take), take);   // :215-216
    }
  }
}

// ------------------------------------------------------------------------------------------
// Huffman baseline
// ------------------------------------------------------------------------------------------
// CPython heapq on node indices ordered by freq only (huffman.py:21-22): exact replica of
// heappush/_siftdown and heappop/_siftup so that ties resolve as in the reference.
__device__ __forceinline__ bool huf_lt(const HufNode* nd, int a, int b) { return nd[a].freq < nd[b].freq; }
__device__ void huf_siftdown(int* heap, const HufNode* nd, int startpos, int pos) {
  const int newitem = heap[pos];
  while (pos > startpos) {
    const int parentpos = (pos - 1) >> 1;
    const int parent = heap[parentpos];
    if (huf_lt(nd, newitem, parent)) { heap[pos] = parent; pos = parentpos; continue; }
    break;
  }
  heap[pos] = newitem;
}
__device__ void huf_siftup(int* heap, const HufNode* nd, int len, int pos) {
  const int startpos = pos;
  const int newitem = heap[pos];
  int childpos = 2 * pos + 1;
  while (childpos < len) {
    const int rightpos = childpos + 1;
    if (rightpos < len && !huf_lt(nd, heap[childpos], heap[rightpos])) childpos = rightpos;
    heap[pos] = heap[childpos];
    pos = childpos;
    childpos = 2 * pos + 1;
  }
  heap[pos] = newitem;
  huf_siftdown(heap, nd, startpos, pos);
}
__device__ int huf_pop(int* heap, const HufNode* nd, int* len) {
  const int last = heap[--(*len)];
  if (*len > 0) {
    const int ret = heap[0];
    heap[0] = last;
    huf_siftup(heap, nd, *len, 0);
    return ret;
  }
  return last;
}

template <bool DECODE>
__global__ void __launch_bounds__(NT, 1) huffman_kernel(ns_codec_params P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CodecShared sm = carve(smem_raw);
  const int row = blockIdx.x, tid = threadIdx.x, V = P.V;
  int slot;
  if (!stream_live(P, row, DECODE, &slot)) return;
  u64 pmax, pmin;
  const float* keys = stage_row_bulk(P, row, sm, true, &pmax, &pmin);
  const float m = key_of_pack(pmax);
  int n = 1 << P.param;                                      // top 2^bits_per_word options, :30
  if (n > V) n = V;
  // log_softmax over the whole row (:32), accumulated in double like torch's CPU kernel
  // (the same sweep fills the count histogram of the selection: staging left it zeroed)
  double acc = 0.0;
  {
    constexpr int NB = HIST_BYTES / 4;
    const float span = m - key_of_pack(pmin);
    const float scale = span > 0.0f ? (float)NB / span : 0.0f;
    for (int i = tid; i < V; i += NT) {
      const float k = keys[i];
      acc += (double)expf(k - m);
      atomicAdd(&sm.hist[bin_of(k, m, scale, NB)], 1u);
    }
  }
  __syncthreads();
  const double sum = block_sum_d(acc, sm.sc->red);
  const float lse = (float)log(sum);
  // the n-th element of the order bounds the kept set
  const int bound_id = element_from_hist(keys, V, pmax, pmin, (u64)(n - 1), sm, P.status ? &P.status[row] : nullptr);
  const u64 bound = bound_id >= 0 ? pack_of(keys[bound_id], bound_id) : pmin;
  const float bkey = key_of_pack(bound);
  const int bid = id_of_pack(bound);
  __syncthreads();
  if (tid == 0) sm.sc->list_count = 0;
  __syncthreads();
  for (int i = tid; i < V; i += NT) {
    const float k = keys[i];
    if (k > bkey || (k == bkey && i <= bid)) {               // pack_of(k, i) >= bound (keys are -0-free)
      const int s = atomicAdd(&sm.sc->list_count, 1);
      if (s < CODEC_LIST_CAP) { sm.list[s].pack = pack_of(k, i); sm.list[s].w = 0; }
    }
  }
  __syncthreads();
  // order the n survivors: w := rank
  for (int c = tid; c < n; c += NT) {
    const u64 pc = sm.list[c].pack;
    int r = 0;
    for (int o = 0; o < n; ++o) r += sm.list[o].pack > pc ? 1 : 0;
    sm.list[c].w = (u64)r;
  }
  __syncthreads();
  // scratch carved from the histogram area (8 KB): ids by rank, nodes, heap
  int* ids = reinterpret_cast<int*>(sm.hist);                          // [n]
  int* heap = ids + HUF_MAX_LEAVES;                                    // [n]
  HufNode* nd = reinterpret_cast<HufNode*>(sm.raw);                    // [2n-1] in the (unused) exp-table + start of hist? no: table area only
  // table area is 4 KB = 256 nodes; larger trees continue in the key row's tail padding -> keep n <= 128 in the table,
  // otherwise place nodes after the list (list holds n <= 512 entries of 16 B = 8 KB, nodes need 16 KB): use global-free fallback
  if (n > 128) nd = reinterpret_cast<HufNode*>(sm.keys + ((V + 8 + 3) & ~3));   // beyond the row (capacity checked on host)
  for (int c = tid; c < n; c += NT) {
    const int r = (int)sm.list[c].w;
    const u64 pc = sm.list[c].pack;
    ids[r] = id_of_pack(pc);
    nd[r].freq = expf((key_of_pack(pc) - m) - lse);                     // probs = exp(log_probs), :33
    nd[r].left = -1; nd[r].right = -1; nd[r].parent = -1;
  }
  __syncthreads();
  if (tid == 0) {
    int len = 0;
    for (int i = 0; i < n; ++i) { heap[len] = i; ++len; huf_siftdown(heap, nd, 0, len - 1); }   // make_heap_from_array
    int next = n;
    while (len > 1) {                                                  // merge_nodes, huffman.py:48-57
      const int n1 = huf_pop(heap, nd, &len);
      const int n2 = huf_pop(heap, nd, &len);
      nd[next].freq = nd[n1].freq + nd[n2].freq;
      nd[next].left = n1; nd[next].right = n2; nd[next].parent = -1;
      nd[n1].parent = next; nd[n2].parent = next;
      heap[len] = next; ++len; huf_siftdown(heap, nd, 0, len - 1);
      ++next;
    }
    const int root = heap[0];
    if (!DECODE) {
      const int cursor = P.cursor[row], mlen = P.msg_len[row];
      const uint32_t* msg = P.msg + (size_t)row * P.msg_stride;
      int node = root, i = cursor;
      while (nd[node].left >= 0) {                                     // :47-52, exhausted message reads as 0
        const int bit = (i < mlen) ? (int)((msg[i >> 5] >> (31 - (i & 31))) & 1u) : 0;
        node = bit ? nd[node].right : nd[node].left;
        ++i;
      }
      emit_token(P, row, slot, ids[node], i - cursor);
    } else {
      int tok = P.token_in[(size_t)row * P.token_stride + slot];
      int leaf = -1;
      for (int r = 0; r < n; ++r) if (ids[r] == tok) leaf = r;
      if (leaf < 0) { leaf = 0; if (P.status) atomicOr(&P.status[row], NS_ST_OUT_OF_RANGE); }   // :149
      u64 code = 0;
      int depth = 0;
      for (int node = leaf; nd[node].parent >= 0; node = nd[node].parent) {
        const int par = nd[node].parent;
        code |= (u64)(nd[par].right == node ? 1 : 0) << depth;          // bits from leaf to root
        ++depth;
      }
      emit_bits(P, row, slot, code, depth);                            // root-to-leaf order = MSB first
    }
  }
}

// ------------------------------------------------------------------------------------------
// bins baseline
// ------------------------------------------------------------------------------------------
template <bool DECODE>
__global__ void __launch_bounds__(NT, 1) bins_kernel(ns_codec_params P) {
  __shared__ u64 red[NWARPS];
  const int row = blockIdx.x, tid = threadIdx.x, V = P.V;
  int slot;
  if (!stream_live(P, row, DECODE, &slot)) return;
  const int b = P.param;
  if (DECODE) {
    if (tid == 0) {
      int tok = P.token_in[(size_t)row * P.token_stride + slot];
      if (tok < 0 || tok >= V) tok = 0;
      const u64 bin = (u64)P.lut[tok];                                 // words2bin[inp[i]], :122
      u64 rev = 0;                                                     // int2bits is LSB first, :183
      for (int j = 0; j < b; ++j) rev |= ((bin >> j) & 1ull) << (b - 1 - j);
      emit_bits(P, row, slot, rev, b);
    }
    return;
  }
  const int cursor = P.cursor[row], mlen = P.msg_len[row];
  const uint32_t* msg = P.msg + (size_t)row * P.msg_stride;
  int bin = 0;
  for (int j = 0; j < b; ++j) {                                        // bits2int(message[i:i+b]), :79
    const int i = cursor + j;
    if (i < mlen) bin |= (int)((msg[i >> 5] >> (31 - (i & 31))) & 1u) << j;
  }
  const float* g = P.logits + (size_t)row * (size_t)P.ld;
  u64 best = 0;
  for (int i = tid; i < V; i += NT) {
    if (P.lut[i] == bin) {
      float k = g[i] + 0.0f;
      if (i == P.mask_id[0] || i == P.mask_id[1]) k = -1e10f;           // :46-47
      const u64 p = pack_of(k, i);
      best = p > best ? p : best;
    }
  }
  best = block_reduce_u(best, OpMaxU(), red);                          // indices[0] of the bin, :80-81
  if (tid == 0) emit_token(P, row, slot, id_of_pack(best), b);         // i += block_size, :85
}


// Throughput form of the bins encoder: one persistent CTA per SM = 16 consumer warps + 1 producer warp.
// The word->bin table is staged once per CTA in shared memory as 16-bit entries; the logits rows stream
// through a ring of bulk-copy slots (full/empty mbarriers, no CTA-wide barrier per slot), so the copy engine
// keeps BINS_RING x 56 KB in flight per SM across row boundaries, independent of warp scheduling.
// HBM-bound: 4*V bytes per token, ~6 instructions per element.
constexpr int BT = 512;                   // consumer threads per CTA
constexpr int BINS_RING = 2;              // slots in flight (2 x 56 KB)
constexpr int BINS_CPT = 7;               // float4 chunks per consumer thread and slot
constexpr int BINS_PC = BINS_CPT * BT;    // float4 chunks per slot (56 KB)
struct BinsRow { int row, slot, bin, cursor, mlen; };
__device__ __forceinline__ void c_mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(c_saddr(bar)) : "memory");
}
__global__ void __launch_bounds__(BT + 32, 1) bins_stream_kernel(ns_codec_params P, int lut_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ u64 red[BT / 32];
  __shared__ u64 full[BINS_RING], empty[BINS_RING];
  __shared__ BinsRow meta[BT];
  __shared__ int wcount[BT / 32];
  __shared__ int n_live;
  uint16_t* lut = reinterpret_cast<uint16_t*>(smem_raw);
  float4* ring = reinterpret_cast<float4*>(smem_raw + lut_bytes);
  const int tid = threadIdx.x, V = P.V, b = P.param;
  const bool producer = tid >= BT;
  for (int i = tid; i < V; i += BT + 32) lut[i] = (uint16_t)P.lut[i];
  if (tid == 0) {
    for (int k = 0; k < BINS_RING; ++k) { c_mbar_init(&full[k], 1); c_mbar_init(&empty[k], BT / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int mk0 = P.mask_id[0], mk1 = P.mask_id[1];
  const int PPR = ((((3 + V + 3) >> 2) - 2) + BINS_PC - 1) / BINS_PC;   // slots per row (the last may be empty)
  const int my_rows = ((int)blockIdx.x < P.B) ? (P.B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  int filled = 0;                                                       // producer: slots issued so far (all batches)
  // chunk count of slot k of a row, and where it starts
  auto geometry = [&](int row, int k, const char** src, int* mis_out) -> int {
    const float* g = P.logits + (size_t)row * (size_t)P.ld;
    const int mis = (int)(((uintptr_t)g & 15u) >> 2);
    const int NI = ((mis + V + 3) >> 2) - 2;                            // interior chunks 1 .. W4-2
    int n = NI - k * BINS_PC;
    if (n > BINS_PC) n = BINS_PC;
    *src = reinterpret_cast<const char*>(g - mis) + 16 + (size_t)k * BINS_PC * 16;
    *mis_out = mis;
    return n > 0 ? n : 0;
  };
  for (int base = 0; base < my_rows; base += BT) {
    // per-row scalars of up to BT rows at once, one consumer thread each: their global latencies overlap and
    // none of them sits on a row's path.  Live rows are compacted in order.
    __syncthreads();
    BinsRow m;
    bool live = false;
    if (!producer && base + tid < my_rows) {
      m.row = blockIdx.x + (base + tid) * gridDim.x;
      const uint8_t phase = P.phase ? P.phase[m.row] : (uint8_t)NS_PHASE_CODING;
      m.slot = P.ntok ? P.ntok[m.row] : 0;
      live = phase != NS_PHASE_DONE;
      if (live && P.ntok && m.slot >= P.token_cap) {
        live = false;
        if (P.phase) P.phase[m.row] = NS_PHASE_DONE;
        if (P.status) atomicOr(&P.status[m.row], NS_ST_TOKEN_OVERFLOW);
      }
      m.cursor = P.cursor[m.row]; m.mlen = P.msg_len[m.row];
      const uint32_t* msg = P.msg + (size_t)m.row * P.msg_stride;
      m.bin = 0;
      for (int j = 0; j < b; ++j) {                                    // bits2int(message[i:i+b]), :79
        const int i = m.cursor + j;
        if (i < m.mlen) m.bin |= (int)((msg[i >> 5] >> (31 - (i & 31))) & 1u) << j;
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, live);
    if (!producer && (tid & 31) == 0) wcount[tid >> 5] = __popc(bal);
    __syncthreads();
    if (!producer) {
      int off = 0;
      for (int w = 0; w < (tid >> 5); ++w) off += wcount[w];
      if (live) meta[off + __popc(bal & ((1u << (tid & 31)) - 1u))] = m;
      if (tid == BT - 1) n_live = off + __popc(bal);
    }
    __syncthreads();
    const int Q = n_live * PPR;                                        // slots of this batch of rows

    if (producer) {
      // ---------------------------------------------------------- producer warp: one lane feeds the ring
      if (tid == BT) {
        for (int q = 0; q < Q; ++q) {
          const char* src; int mis;
          const int n = geometry(meta[q / PPR].row, q % PPR, &src, &mis);
          if (n == 0) continue;                                        // consumers skip it too
          const int s = filled % BINS_RING;
          if (filled >= BINS_RING) c_mbar_wait(&empty[s], ((filled / BINS_RING) - 1) & 1u);   // slot drained
          c_mbar_expect_tx(&full[s], (uint32_t)n * 16u);
          c_bulk_g2s(ring + (size_t)s * BINS_PC, src, (uint32_t)n * 16u, &full[s]);
          ++filled;
        }
      }
    } else {
      // ---------------------------------------------------------- consumers
      float bk = __int_as_float(0x7fc00000), ek = 0.0f;                 // best key of this thread (lowest id among equals)
      int bi = 0x7fffffff, eid = -1;
      uint16_t bin = 0;
      int rr = 0, k = 0, row = 0, mis = 0, NI = 0;
      for (int q = 0; q < Q; ++q, ++k) {
        if (k == PPR) { k = 0; ++rr; }
        if (k == 0) {                                                  // a new row: its geometry and scalars
          row = meta[rr].row;
          bin = (uint16_t)meta[rr].bin;
          const float* g = P.logits + (size_t)row * (size_t)P.ld;
          mis = (int)(((uintptr_t)g & 15u) >> 2);
          const int W4 = (mis + V + 3) >> 2;
          NI = W4 - 2;
          bk = __int_as_float(0x7fc00000); bi = 0x7fffffff; eid = -1;
          if (tid < 8) {                                               // edge chunks: element-wise, used at the row end
            const int id = 4 * (tid < 4 ? 0 : W4 - 1) - mis + (tid & 3);
            if (id >= 0 && id < V) { eid = id; ek = g[id]; }
          }
        }
        int n = NI - k * BINS_PC;
        if (n > BINS_PC) n = BINS_PC;
        if (n > 0) {
          const int s = filled % BINS_RING;
          c_mbar_wait(&full[s], (filled / BINS_RING) & 1u);
          ++filled;
          const float4* piece = ring + (size_t)s * BINS_PC;
          const int c_base = 1 + k * BINS_PC;
          const int cm0 = (mk0 >= 0 && mk0 < V) ? ((mk0 + mis) >> 2) - c_base : -1;   // chunk of a forbidden token
          const int cm1 = (mk1 >= 0 && mk1 < V) ? ((mk1 + mis) >> 2) - c_base : -1;
#pragma unroll
          for (int j = 0; j < BINS_CPT; ++j) {
            const int c = tid + j * BT;
            if (c < n) {
              float4 v = piece[c];
              const int id = 4 * (c_base + c) - mis;
              if (c == cm0 || c == cm1) {                              // rare: forbidden tokens, :46-47
                float* f = reinterpret_cast<float*>(&v);
#pragma unroll
                for (int e = 0; e < 4; ++e) if (id + e == mk0 || id + e == mk1) f[e] = -1e10f;
              }
              // bk starts as NaN: "not (x <= bk)" takes the first member of the bin, then only strictly larger
              // keys (ids ascend within a thread; -0 == +0 here and is folded when the key is packed)
              if (lut[id] == bin && !(v.x <= bk)) { bk = v.x; bi = id; }
              if (lut[id + 1] == bin && !(v.y <= bk)) { bk = v.y; bi = id + 1; }
              if (lut[id + 2] == bin && !(v.z <= bk)) { bk = v.z; bi = id + 2; }
              if (lut[id + 3] == bin && !(v.w <= bk)) { bk = v.w; bi = id + 3; }
            }
          }
          __syncwarp();
          if ((tid & 31) == 0) c_mbar_arrive(&empty[s]);               // this warp is done with the slot
        }
        if (k == PPR - 1) {                                            // row complete: merge through the (key, id) pack order
          u64 best = (bi != 0x7fffffff) ? pack_of(bk + 0.0f, bi) : 0ull;
          if (eid >= 0 && lut[eid] == bin) {
            float kk = ek + 0.0f;
            if (eid == mk0 || eid == mk1) kk = -1e10f;
            const u64 p = pack_of(kk, eid);
            best = p > best ? p : best;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) { const u64 x = __shfl_xor_sync(0xffffffffu, best, o); best = x > best ? x : best; }
          asm volatile("bar.sync 1, %0;" :: "n"(BT) : "memory");       // consumers only: `red` of the previous row is consumed
          if ((tid & 31) == 0) red[tid >> 5] = best;
          asm volatile("bar.sync 1, %0;" :: "n"(BT) : "memory");
          if (tid == 0) {
            u64 w = red[0];
            for (int i = 1; i < BT / 32; ++i) w = red[i] > w ? red[i] : w;
            const BinsRow mr = meta[rr];
            // indices[0] of the bin, :80-81; i += block_size, :85
            const int nc = mr.cursor + b;
            P.cursor[row] = nc;
            P.token_out[(size_t)row * P.token_stride + mr.slot] = id_of_pack(w);
            if (P.ntok) P.ntok[row] = mr.slot + 1;
            if (P.nbits_out) P.nbits_out[row] = (uint8_t)b;
            if (P.phase && nc >= mr.mlen) P.phase[row] = NS_PHASE_DONE;
          }
        }
      }
    }
  }
}

thread_local char g_cerr[256] = "";
int cerr(int code, const char* msg) { snprintf(g_cerr, sizeof(g_cerr), "%s", msg); return code; }

int validate_codec(const ns_codec_params* p, int kind) {
  if (!p) return cerr(NS_E_NULL, "params is NULL");
  const bool decode = (kind == K_RANK_DEC || kind == K_HUF_DEC || kind == K_BINS_DEC);
  const bool needs_logits = kind != K_BINS_DEC;
  if (needs_logits && !p->logits) return cerr(NS_E_NULL, "logits is NULL");
  if (p->B < 0 || p->V < 4) return cerr(NS_E_RANGE, "B < 0 or V < 4");
  if (needs_logits && p->ld < p->V) return cerr(NS_E_RANGE, "ld < V");
  if ((kind == K_RANK_ENC || kind == K_RANK_DEC || kind == K_HUF_ENC || kind == K_HUF_DEC) && p->V > MAX_VOCAB)
    return cerr(NS_E_VOCAB, "V exceeds the shared-memory row capacity");
  if (kind == K_RANK_ENC || kind == K_RANK_DEC) { if (!(p->temp > 0.0)) return cerr(NS_E_RANGE, "temp must be > 0"); }
  if (kind == K_HUF_ENC || kind == K_HUF_DEC) { if (p->param < 1 || p->param > 9) return cerr(NS_E_RANGE, "bits_per_word must be in [1, 9]"); }
  if (kind == K_BINS_ENC || kind == K_BINS_DEC) {
    if (p->param < 1 || p->param > 16) return cerr(NS_E_RANGE, "block_size must be in [1, 16]");
    if (!p->lut) return cerr(NS_E_NULL, "bins need the word->bin table");
  }
  if (!decode) { if (!p->msg || !p->msg_len || !p->cursor || !p->token_out) return cerr(NS_E_NULL, "encode needs msg/msg_len/cursor/token_out"); }
  else { if (!p->token_in || !p->out_bits || !p->out_len) return cerr(NS_E_NULL, "decode needs token_in/out_bits/out_len"); }
  return NS_OK;
}

#include "ns_codecs_stream.cuh"

// row-streaming kernels: one small CTA per row, several rows per SM
template <typename K>
int launch_stream_codec(K kernel, const ns_codec_params* p, void* stream) {
  if (p->B == 0) return NS_OK;
  kernel<<<p->B, CT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cerr((int)e, cudaGetErrorString(e));
  return NS_OK;
}
// rows they serve: every lane group owns elements (the Huffman bound needs 32 group maxima)
bool stream_ok(const ns_codec_params* p) { return p->V >= 4 * CT && getenv("NS_CODEC_STREAM_OFF") == nullptr; }

template <typename K>
int launch_codec(K kernel, const ns_codec_params* p, int smem, bool* configured, void* stream) {
  if (!*configured && smem > 0) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) { cerr((int)e, cudaGetErrorString(e)); return e == cudaErrorInvalidDeviceFunction ? NS_E_NODEVICE : (int)e; }
    *configured = true;
  }
  if (p->B == 0) return NS_OK;
  kernel<<<p->B, NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(*p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cerr((int)e, cudaGetErrorString(e));
  return NS_OK;
}

int row_smem(const ns_codec_params* p, bool huffman) {
  int bytes = FIXED_BYTES + (p->V + 8) * 4;
  if (huffman) bytes += 2 * HUF_MAX_LEAVES * (int)sizeof(HufNode) + 64;   // node pool behind the row for large trees
  return bytes;
}

}  // namespace

extern "C" {

int ns_sizeof_codec_params(void) { return (int)sizeof(ns_codec_params); }

const char* ns_codec_last_error_string(void) { return g_cerr; }

int ns_rank_encode_step(const ns_codec_params* p, void* s) {
  int rc = validate_codec(p, K_RANK_ENC); if (rc) return rc;
  if (stream_ok(p) && !(p->top_p > 0.0 && p->top_p < 1.0) && !(p->min_prob > 0.0)) return launch_stream_codec(codec_stream_kernel<K_RANK_ENC>, p, s);
  static bool c = false; return launch_codec(rank_kernel<false>, p, row_smem(p, false), &c, s);
}
int ns_rank_decode_step(const ns_codec_params* p, void* s) {
  int rc = validate_codec(p, K_RANK_DEC); if (rc) return rc;
  if (stream_ok(p) && !(p->top_p > 0.0 && p->top_p < 1.0) && !(p->min_prob > 0.0)) return launch_stream_codec(codec_stream_kernel<K_RANK_DEC>, p, s);
  static bool c = false; return launch_codec(rank_kernel<true>, p, row_smem(p, false), &c, s);
}
int ns_huffman_encode_step(const ns_codec_params* p, void* s) {
  int rc = validate_codec(p, K_HUF_ENC); if (rc) return rc;
  if (stream_ok(p) && p->param <= 5) return launch_stream_codec(codec_stream_kernel<K_HUF_ENC>, p, s);
  if (row_smem(p, true) > SMEM_LIMIT && p->param > 7) return cerr(NS_E_VOCAB, "bits_per_word > 7 needs a smaller vocabulary");
  static bool c = false; return launch_codec(huffman_kernel<false>, p, row_smem(p, p->param > 7), &c, s);
}
int ns_huffman_decode_step(const ns_codec_params* p, void* s) {
  int rc = validate_codec(p, K_HUF_DEC); if (rc) return rc;
  if (stream_ok(p) && p->param <= 5) return launch_stream_codec(codec_stream_kernel<K_HUF_DEC>, p, s);
  if (row_smem(p, true) > SMEM_LIMIT && p->param > 7) return cerr(NS_E_VOCAB, "bits_per_word > 7 needs a smaller vocabulary");
  static bool c = false; return launch_codec(huffman_kernel<true>, p, row_smem(p, p->param > 7), &c, s);
}
int ns_bins_encode_step(const ns_codec_params* p, void* s) {
  int rc = validate_codec(p, K_BINS_ENC); if (rc) return rc;
  const int lut_bytes = ((p->V * 2) + 127) & ~127;
  const int dyn = lut_bytes + BINS_RING * BINS_PC * 16;
  if (dyn + 16 * 1024 <= SMEM_LIMIT && p->B > 0) {                   // table + ring fit: streaming kernel
    static bool c = false;
    static int sms = 0;
    if (!c) {
      cudaError_t e = cudaFuncSetAttribute(bins_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT - 16 * 1024);
      if (e != cudaSuccess) { cerr((int)e, cudaGetErrorString(e)); return e == cudaErrorInvalidDeviceFunction ? NS_E_NODEVICE : (int)e; }
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      c = true;
    }
    int grid = sms;
    if (grid > p->B) grid = p->B;
    bins_stream_kernel<<<grid, BT + 32, dyn, reinterpret_cast<cudaStream_t>(s)>>>(*p, lut_bytes);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cerr((int)e, cudaGetErrorString(e));
    return NS_OK;
  }
  static bool c = false; return launch_codec(bins_kernel<false>, p, 0, &c, s);
}
int ns_bins_decode_step(const ns_codec_params* p, void* s) {
  int rc = validate_codec(p, K_BINS_DEC); if (rc) return rc;
  static bool c = false; return launch_codec(bins_kernel<true>, p, 0, &c, s);
}

}  // extern "C"
