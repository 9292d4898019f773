/*
 * ns_coder.h -- C ABI of the B200-native steganographic coder step.
 *
 * The reference (nobkagit/NeuralSteganography) is pure Python and has no FFI;
 * its boundary for this path is the Python LMProvider protocol
 * (src/neuralstego/api.py:42-56).  This header is the native boundary a
 * maintainer binds with ctypes (INTEGRATION.md shows the stub): plain
 * pointers and sizes, no torch types.  Every entry point
 *   - takes DEVICE pointers (unless the name ends in _host) and a cudaStream_t
 *     passed as void*,
 *   - never allocates, never synchronises, is CUDA-graph capturable,
 *   - returns 0 on success, a negative NS_E_* code for argument errors, or a
 *     positive cudaError_t.
 *
 * One "stream" is one independent message/cover pair (one chunk of
 * stego_encode, api.py:736-747).  All per-stream arrays have B entries.
 */
#ifndef NS_CODER_H
#define NS_CODER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NS_ABI_VERSION 1

/* error codes */
#define NS_OK 0
#define NS_E_NULL (-1)        /* required pointer is NULL */
#define NS_E_RANGE (-2)       /* precision/topk/temp/B/V out of range */
#define NS_E_VOCAB (-3)       /* vocabulary does not fit the kernel's shared-memory row */
#define NS_E_ALIGN (-4)       /* pointer not 4-byte aligned */
#define NS_E_NODEVICE (-5)    /* no sm_100 device / kernel image missing */

/* per-stream phase (uint8) */
#define NS_PHASE_CODING 0     /* message bits remain: full arithmetic step            */
#define NS_PHASE_TAIL 1       /* finish_sent tail: emit rank-0 token, no interval update
                                 (code_base/arithmetic.py:135-137)                    */
#define NS_PHASE_DONE 2       /* stream finished: kernel leaves it untouched          */

/* per-stream status bits written by the kernels (int32, OR-ed, 0 = fine) */
#define NS_ST_OUT_OF_RANGE 1  /* decode: observed token is not in the kept set (rank >= k,
                                 code_base/arithmetic.py:301) -- coded as rank 0 like :342 */
#define NS_ST_BIN_OVERFLOW 2  /* selection bucket exceeded the resolve capacity (degenerate
                                 logits, e.g. thousands of exactly equal values)       */
#define NS_ST_EST_RETRY 4     /* informational: the throughput kernel handed this row to the exact
                                 multi-pass kernel (top-k inside the cutoff set, estimate outside
                                 its guard band, list overflow); results are identical */
#define NS_ST_TOKEN_OVERFLOW 8 /* encode: token buffer full (ntok >= token_cap); stream stopped */
#define NS_ST_RANK_DEFER 16   /* informational: the sweep kernel of the rank form (rank_ws given) left this row to the
                                 row-resident kernel (not certainly top-k bound, finish_sent tail); results are identical */

/* Arithmetic coder (A): code_base/arithmetic.py:78-217 (encode), :220-373 (decode). */
typedef struct ns_ac_params {
  /* logits: fp32 [B, V], row r at logits + r*ld (elements).  Read-only: the
     reference's in-place masking (arithmetic.py:124-125) is applied on chip. */
  const float* logits;
  int64_t ld;
  int32_t B;
  int32_t V;
  double temp;          /* arithmetic.py:85,129  */
  int32_t precision;    /* arithmetic.py:86,96   (1..48) */
  int32_t topk;         /* arithmetic.py:87,142  */
  int32_t mask_id[2];   /* forbidden token ids (arithmetic.py:124-125: V-1 and 628); -1 = none */
  /* coder state, arithmetic.py:98: cur_interval = [lo, hi) */
  uint64_t* lo;
  uint64_t* hi;
  uint8_t* phase;       /* NS_PHASE_*; may be NULL (= all coding) */
  int32_t* status;      /* NS_ST_* bits, OR-ed in; may be NULL */
  /* Per-stream token counter.  When non-NULL the step reads/writes token slot
     ntok[r] of its stream and increments it, so the same parameter block (and a
     captured CUDA graph) serves every step of the generation loop; when NULL the
     slot is 0.  Encode stops a stream whose counter reaches token_cap; decode
     marks a stream DONE when the counter reaches ntok_total[r] and treats its
     last token as arithmetic.py:356 does. */
  int32_t* ntok;
  int32_t token_cap;
  const int32_t* ntok_total;
  /* ---- encode only ---------------------------------------------------- */
  /* message bits packed MSB-first: bit t of stream r is
     (msg[r*msg_stride + (t>>5)] >> (31 - (t&31))) & 1   (arithmetic.py:168-171) */
  const uint32_t* msg;
  int64_t msg_stride;   /* words per stream */
  const int32_t* msg_len;   /* bits per stream */
  int32_t* cursor;      /* "i" of arithmetic.py:112,184 */
  int32_t* token_out;   /* token chosen this step: token_out[r*token_stride + slot] */
  int64_t token_stride;
  int32_t finish_sent;  /* arithmetic.py:83,114 */
  const uint8_t* sent_end;  /* [V] 1 if the token text contains . ! ? (utils.py:55-57); may be NULL */
  /* ---- decode only ---------------------------------------------------- */
  const int32_t* token_in;  /* observed token: token_in[r*token_stride + slot] */
  const uint8_t* is_last;   /* [B] 1 on the stream's last token (arithmetic.py:356); may be NULL;
                               ignored when ntok_total is given */
  uint32_t* out_bits;   /* recovered bits, packed MSB-first like msg */
  int64_t out_stride;   /* words per stream */
  int32_t* out_len;     /* bits written so far per stream */
  /* ---- both ----------------------------------------------------------- */
  uint8_t* nbits_out;   /* bits consumed/emitted this step (arithmetic.py:183); may be NULL */
  /* optional per-step trace for parity tests: [B,4] = new_bottom, new_top, k, total mass */
  uint64_t* trace;
  /* Work queue, B+2 int32 zeroed once by the caller.  When given (and precision <= 31) a step is
     two launches: the single-pass throughput kernel, then the exact multi-pass kernel on the rows
     the first one queued here (top-k inside the cutoff set, degenerate rows).  NULL = exact kernel
     only.  force_exact != 0 also selects the exact kernel only. */
  int32_t* slow_ws;
  int32_t force_exact;
  /* optional profiling counters (16 x uint64, zeroed by the caller): SM-clock cycles spent per
     phase of the throughput kernel, summed over thread 0 of every CTA; [15] counts rows */
  uint64_t* prof;
  /* optional, encode only: per-step statistics of the reference (arithmetic.py:192-198), [B,3] fp64 =
     log p(selected token), KL(q_hat || p) in bits, entropy of the tempered distribution in bits.
     Requesting them routes the step through the exact kernel. */
  double* stats;
  /* Optional second work queue, B+2 int32 zeroed once by the caller (same layout as slow_ws).  When given, a step whose
     top-k binds (2 <= topk <= 512, V >= 8192 and V >= 32 (2.5 topk + 64)) starts with the sweep kernel of the rank form
     (csrc/ns_topk.cuh: the row is read once, no resident row, four rows per SM); rows it does not carry are queued here
     and done by the row-resident kernel (csrc/ns_fast.cuh), which also serves the smaller vocabularies.
     NULL = row-resident kernel for every row.  Results are identical. */
  int32_t* rank_ws;
  /* Reserved; must be 0. */
  int64_t scratch_stride;
  int32_t scratch_slots;
  /* Kernel choice for the throughput path: 0 = default (threshold form of the cutoff: the lean single-row kernel
     ns_lean.cuh; rank form, topk <= 512: ns_topk.cuh when rank_ws is given, then ns_fast.cuh), 2 = always ns_fast.cuh.
     Results are identical. */
  int32_t variant;
} ns_ac_params;

int ns_version(void);
const char* ns_last_error_string(void);
/* sizeof(ns_ac_params) as compiled, so a binding can verify its struct mirror */
int ns_sizeof_ac_params(void);
/* largest V the arithmetic-coder kernels accept on this build (the row lives in shared memory) */
int ns_ac_max_vocab(void);

/* one encode step for B streams (code_base/arithmetic.py:114-210 loop body) */
int ns_ac_encode_step(const ns_ac_params* p, void* cuda_stream);
/* one decode step for B streams (code_base/arithmetic.py:255-371 loop body) */
int ns_ac_decode_step(const ns_ac_params* p, void* cuda_stream);
/* parity-test helper: integer bin widths of the kept set per token id,
   q_out [B,V] uint64 (0 = not kept), meta_out [B,4] = k, slack, total, range.
   (code_base/arithmetic.py:146-158) */
int ns_ac_debug_bins(const ns_ac_params* p, uint64_t* q_out, uint64_t* meta_out, void* cuda_stream);

/* ------------------------------------------------------------------------------------------
 * Comparison codecs of the reference, same batching and token bookkeeping as ns_ac_params:
 *   rank    src/neuralstego/codec/arithmetic.py:122-231 (encode_with_lm / decode_with_lm),
 *           :370-385 (_rank_tokens); temperature as in lm/arithmetic.py:69-73
 *   huffman code_base/huffman_baseline.py:7-71, :73-165; code_base/huffman.py:12-76
 *   bins    code_base/block_baseline.py:9-24, :26-97, :99-189
 * Message / output bits are packed like ns_ac_params.msg: bit t of a stream is the t-th element
 * of the reference's Python bit list.
 * ------------------------------------------------------------------------------------------ */
typedef struct ns_codec_params {
  const float* logits;      /* fp32 [B, V], row stride ld */
  int64_t ld;
  int32_t B;
  int32_t V;
  double temp;              /* rank: temperature of the provider (lm/arithmetic.py:72) */
  int32_t param;            /* huffman: bits_per_word (1..9); bins: block_size; rank: unused */
  int32_t topk;             /* rank: quality top_k (quality.py:76-81); <= 0 = none */
  int32_t mask_id[2];       /* huffman/bins: tokens forced to -1e10 (huffman_baseline.py:26-27); -1 = none */
  uint8_t* phase;           /* NS_PHASE_CODING / NS_PHASE_DONE per stream; may be NULL */
  int32_t* status;          /* NS_ST_* bits; may be NULL */
  int32_t* ntok;            /* per-stream token slot, as in ns_ac_params */
  int32_t token_cap;
  const int32_t* ntok_total;
  const uint32_t* msg;      /* encode */
  int64_t msg_stride;
  const int32_t* msg_len;
  int32_t* cursor;
  int32_t* token_out;
  int64_t token_stride;
  const int32_t* token_in;  /* decode */
  uint32_t* out_bits;
  int64_t out_stride;
  int32_t* out_len;
  const int32_t* total_bits; /* rank decode: payload bit count (state["residual_bits"], arithmetic.py:167) */
  uint8_t* nbits_out;       /* bits consumed / emitted this step; may be NULL */
  const int32_t* lut;       /* bins: word -> bin [V] (get_bins, block_baseline.py:9-24) */
  double top_p;             /* rank: quality top_p (quality.py:85-91), used when 0 < top_p < 1; else none */
  double min_prob;          /* rank: quality min_prob (quality.py:93-96), used when > 0; else none */
} ns_codec_params;

int ns_sizeof_codec_params(void);
const char* ns_codec_last_error_string(void);
/* (B) rank codec: floor(log2(#tokens with p>0)) message bits pick the token of that rank */
int ns_rank_encode_step(const ns_codec_params* p, void* cuda_stream);
int ns_rank_decode_step(const ns_codec_params* p, void* cuda_stream);
/* Huffman baseline over the top 2^bits_per_word tokens (heapq-compatible tree) */
int ns_huffman_encode_step(const ns_codec_params* p, void* cuda_stream);
int ns_huffman_decode_step(const ns_codec_params* p, void* cuda_stream);
/* bins baseline: block_size message bits pick a vocabulary bin, token = argmax inside it */
int ns_bins_encode_step(const ns_codec_params* p, void* cuda_stream);
int ns_bins_decode_step(const ns_codec_params* p, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* NS_CODER_H */
