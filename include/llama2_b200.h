/*
 * llama2_b200.h -- C ABI of libllama2_b200.so
 *
 * Drop-in boundary for ONE path of wizzard0/llama2.ts: the call
 *     transformer(token, pos, config, state, weights)        llama2.ts:468
 * and the data it consumes/produces (weights llama2.ts:95-129, RunState
 * llama2.ts:131-163, logits read at llama2.ts:478-492).  Everything on the
 * device (weights, activations, KV cache, kernels, graphs) lives behind the
 * opaque l2b_ctx.  All arguments are int32/uint64/pointers so bun:ffi and a
 * Node N-API shim can both bind them.
 *
 * Conventions
 *   - every function returns 0 on success, a negative L2B_E* code on failure;
 *     l2b_last_error() returns the text.  Nothing aborts or throws across the
 *     ABI.
 *   - host buffers are borrowed only for the duration of the call.
 *   - calls on one ctx are not re-entrant (the reference is one JS thread).
 *   - there is NO CPU fallback: without a CUDA device every entry point that
 *     touches the device fails with L2B_ECUDA.
 */
#ifndef LLAMA2_B200_H
#define LLAMA2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define L2B_ABI_VERSION 2

/* error codes */
#define L2B_OK        0
#define L2B_EINVAL   -1   /* bad argument (null pointer, id/shape/range)        */
#define L2B_EORDER   -2   /* pos not callable yet (KV rows < pos never written) */
#define L2B_ECUDA    -3   /* CUDA runtime / driver failure, or no device        */
#define L2B_ESTATE   -4   /* weights not fully uploaded, wrong mode for call    */
#define L2B_ENOMEM   -5   /* device or host allocation failed                   */
#define L2B_ECOMM    -6   /* tensor-parallel exchange failure                   */

/* tensor ids: field order of `interface TransformerWeights`, llama2.ts:95-110 */
#define L2B_T_TOKEN_EMBEDDING_TABLE 0   /* (vocab, dim)            layer = 0   */
#define L2B_T_RMS_ATT_WEIGHT        1   /* (dim)                   per layer   */
#define L2B_T_WQ                    2   /* (dim, dim)              per layer   */
#define L2B_T_WK                    3
#define L2B_T_WV                    4
#define L2B_T_WO                    5
#define L2B_T_RMS_FFN_WEIGHT        6   /* (dim)                   per layer   */
#define L2B_T_W1                    7   /* (hidden, dim)           per layer   */
#define L2B_T_W2                    8   /* (dim, hidden)           per layer   */
#define L2B_T_W3                    9   /* (hidden, dim)           per layer   */
#define L2B_T_RMS_FINAL_WEIGHT      10  /* (dim)                   layer = 0   */
#define L2B_T_FREQ_CIS_REAL         11  /* (seq_len, head_size/2)  layer = 0   */
#define L2B_T_FREQ_CIS_IMAG         12
#define L2B_T_WCLS                  13  /* (vocab, dim); only when hdr[5] < 0  */
#define L2B_T_COUNT                 14

typedef struct l2b_ctx l2b_ctx;

/* Replaces readConfig() + newRunState() (llama2.ts:80-93, 147-163).
 *   hdr        the 7 int32 header words of the checkpoint, unmodified:
 *              dim, hidden_dim, n_layers, n_heads, n_kv_heads, +/-vocab, seq_len
 *              (n_kv_heads is accepted and ignored, like llama2.ts:86; a
 *              negative vocab means an unshared classifier, llama2.ts:87-90).
 *   device     CUDA device ordinal this ctx lives on.
 *   max_batch  number of independent sequences (RunStates) held; 1 = the
 *              reference's batch-1 decode.
 *   max_steps  rows of KV cache per sequence and layer (<= seq_len; 0 = seq_len).
 *              The reference allocates seq_len rows (llama2.ts:160-161) but can
 *              only ever touch `steps` of them (llama2.ts:439).                 */
int l2b_create(const int32_t hdr[7], int32_t device, int32_t max_batch,
               int32_t max_steps, l2b_ctx** out);

/* The boundary exactly as SURVEY.md section 8(b) writes it -- l2b_create(hdr, n_gpus, tp_degree,
 * max_batch, max_steps) -- for the reference's ONE host thread (llama2.ts:468 is a synchronous call
 * from a single JS thread): the calling thread drives CUDA devices 0..n_gpus-1 itself, no
 * second process, no IPC handles, no torch.
 *   tp_degree == 1        the max_batch sequences are partitioned over the GPUs (sequence b lives on
 *                         GPU b / ceil(max_batch / n_gpus)); weights are replicated by l2b_upload /
 *                         l2b_load_checkpoint; no collective touches the data path.
 *   tp_degree == n_gpus   ONE tensor-parallel group decoding one sequence (max_batch must be 1):
 *                         rows of every projection sharded over the GPUs, slices exchanged by peer
 *                         stores over NVLink inside the kernels (cudaDeviceEnablePeerAccess).
 * Every entry point below accepts the returned handle; l2b_tp_export/l2b_tp_connect and
 * l2b_prefill (tensor-parallel group) do not apply.  l2b_create(hdr, device, ...) remains the
 * one-GPU form with an explicit device ordinal; l2b_create_tp is the one-process-per-GPU form. */
int l2b_create_multi(const int32_t hdr[7], int32_t n_gpus, int32_t tp_degree, int32_t max_batch,
                     int32_t max_steps, l2b_ctx** out);

/* Tensor-parallel variant (row-sharded projections): this
 * process is rank `tp_rank` of `tp_size` (one process per GPU).  Heads, hidden
 * rows and vocab rows are split evenly; uploads still pass the FULL tensor and
 * the library keeps its slice.  Exchange buffers are wired with
 * l2b_tp_export/l2b_tp_connect below.
 * RENDEZVOUS: every rank must issue the same sequence of step calls (same token, pos, n_steps),
 * and must enter each call within the exchange time-out of its peers (option "tp_timeout_ms",
 * default 20000): a kernel spins on its peers' data for at most that long, then the call fails
 * with L2B_ECOMM on every rank of the group (the failure is propagated with the step's last
 * exchange), no token/pos state is advanced and the exchange epoch still moves on, so the same
 * step can be re-issued by ALL ranks.  Put a host barrier between loading the weights and the
 * first step (dist.connect_tp does), because load times differ between ranks.                 */
int l2b_create_tp(const int32_t hdr[7], int32_t device, int32_t max_steps,
                  int32_t tp_rank, int32_t tp_size, l2b_ctx** out);

/* Replaces holding the Float32Array slices of readWeights() (llama2.ts:112-129).
 * COPIES n_floats from host into HBM (re-laid out for the kernels); the caller
 * may free `host` on return.  `host` may also be a device pointer whose contents
 * are complete (the copy is cudaMemcpyDefault on the ctx stream).  `layer` is the index into the per-layer arrays
 * (0 for the unlayered tensors).  L2B_T_WCLS must not be uploaded for a shared
 * classifier -- the library aliases the embedding table like llama2.ts:127.     */
int l2b_upload(l2b_ctx* ctx, int32_t tensor_id, int32_t layer, const float* host,
               uint64_t n_floats);

/* Checkpoint loader fast path: reads a llama2.c legacy-v0
 * .bin (the file llama2.ts:427-436 opens) straight into the device layout through pinned
 * double-buffered staging; equivalent to the l2b_upload calls of readWeights()
 * (llama2.ts:112-129) in file order.  The header must match the context.  A tensor-parallel
 * rank reads only its own rows.  seconds_out (may be NULL) receives the wall time.       */
int l2b_load_checkpoint(l2b_ctx* ctx, const char* path, double* seconds_out);

/* 1 when every tensor the config needs has been uploaded, else 0. */
int l2b_weights_ready(const l2b_ctx* ctx);

/* Replaces transformer(token,pos,...) (llama2.ts:205-303, call site :468) for
 * sequence 0: runs one decode step and copies the vocab_size logits to
 * `logits_out` (host), i.e. what state.logits holds at llama2.ts:478-492.
 * pos must be <= the number of positions already run for this sequence.        */
int l2b_forward(l2b_ctx* ctx, int32_t token, int32_t pos, float* logits_out);

/* Same step, but returns argmax(state.logits) (llama2.ts:364-366 semantics:
 * first maximum wins) computed on the device -- the `-t 0` path, llama2.ts:478. */
int l2b_forward_argmax(l2b_ctx* ctx, int32_t token, int32_t pos, int32_t* next_out);

/* Device-side sampling.  One decode step followed by what
 * the host does at llama2.ts:476-494 -- logits /= temperature, softmax, then sample()
 * (topp <= 0 or >= 1, llama2.ts:368-376) or sample_topp() (llama2.ts:378-394, including its
 * exclusive `i < lastIdx` walk and the fallback to token 0) -- with only the chosen token
 * coming back.  rand01 is the host's random_f32() for this token (the reference draws exactly
 * one per sampled token, :370/:388), so the xorshift stream stays on the host, unchanged.
 * temperature == 0 is the argmax path.  temperature/topp are doubles like JS numbers.       */
int l2b_forward_sample(l2b_ctx* ctx, int32_t token, int32_t pos, double temperature, double topp,
                       float rand01, int32_t* next_out);
/* The same sampler on caller-supplied logits (vocab floats, host); for tests and tools.   */
int l2b_sample_logits(l2b_ctx* ctx, const float* logits_host, double temperature, double topp,
                      float rand01, int32_t* next_out);

/* B independent sequences advance one step each (sequence b uses RunState b).
 * tokens[b], pos[b] as above; logits_out (B*vocab floats, host) may be NULL;
 * argmax_out (B ints, host) may be NULL.
 * NUMERICS per path (also l2b_generate_greedy, l2b_profile_batch):
 *   B <  tc_min_batch (default 3), or max_batch < 3: the reference-exact path -- fp32 storage,
 *       fp64 accumulation like JS numbers; the same results as B separate l2b_forward calls.
 *   B >= tc_min_batch on a ctx created with max_batch >= 3: tcgen05 GEMMs with 3xTF32 products,
 *       fp32 accumulation and an fp32 expf in SwiGLU -- inside the 1e-4 abs / 1e-3 rel tolerance
 *       but NOT bit-identical to batch-1 calls, and logits depend slightly on B (k-split choice).
 *       Callers that need bit parity with l2b_forward set option "tc_min_batch" to 0.
 *   On the tensor-core path hb is not materialised: l2b_read_state(L2B_S_HB) fails with L2B_ESTATE. */
int l2b_forward_batch(l2b_ctx* ctx, int32_t B, const int32_t* tokens, const int32_t* pos,
                      float* logits_out, int32_t* argmax_out);

/* Prompt prefill.  The reference feeds the prompt one token
 * at a time through transformer() and throws the logits away (llama2.ts:465-474).  This call
 * has the same effect on sequence `seq` as n_tokens successive l2b_forward(tokens[i],
 * pos0 + i) calls -- KV rows pos0..pos0+n_tokens-1 are written -- but runs all positions in
 * ONE pass over the weights on the tensor cores (causal attention inside the batch).
 * logits_out (vocab floats) / argmax_out receive the LAST position's logits / argmax; either
 * may be NULL.  Numerics: the 3xTF32 path (within the 1e-4 / 1e-3 tolerance, not bit-exact). */
int l2b_prefill(l2b_ctx* ctx, int32_t seq, int32_t n_tokens, const int32_t* tokens, int32_t pos0,
                float* logits_out, int32_t* argmax_out);

/* The greedy generate loop of llama2.ts:465-508 kept on the device: starting
 * from `token` at `pos`, runs `n_steps` steps; step i feeds forced[i] when
 * forced != NULL and forced[i] >= 0 (prompt forcing, llama2.ts:471-473), else
 * the device argmax.  out_tokens[i] receives the `next` of step i.  Does not
 * stop at BOS (the caller truncates, llama2.ts:499).  For B > 1 every array is
 * [n_steps][B] and `tokens`/`pos` have B entries.                              */
int l2b_generate_greedy(l2b_ctx* ctx, int32_t B, const int32_t* tokens, const int32_t* pos,
                        int32_t n_steps, const int32_t* forced, int32_t* out_tokens);

/* Device time (ms, CUDA events on the ctx stream) of the last l2b_forward*,
 * l2b_forward_batch or l2b_generate_greedy call, and the number of kernel
 * launches it issued.                                                           */
float   l2b_last_device_ms(const l2b_ctx* ctx);
int64_t l2b_last_launches(const l2b_ctx* ctx);

/* Kernel classes of one decode step (index into l2b_profile_step's arrays).   */
#define L2B_K_QKV   0   /* rmsnorm -> wq/wk/wv matvec -> RoPE -> KV write  llama2.ts:216-240 */
#define L2B_K_ATTN  1   /* scores / softmax / weighted sum                llama2.ts:244-267 */
#define L2B_K_WO    2   /* wo matvec + residual                           llama2.ts:270-273 */
#define L2B_K_W13   3   /* rmsnorm -> w1/w3 matvec -> SwiGLU              llama2.ts:276-289 */
#define L2B_K_W2    4   /* w2 matvec + residual                           llama2.ts:292-295 */
#define L2B_K_CLS   5   /* final rmsnorm -> wcls matvec -> argmax         llama2.ts:299-302 */
/* batched tensor-core path (B >= tc_min_batch): the same five matmuls as tcgen05 GEMMs   */
#define L2B_K_GEMM_QKV 6
#define L2B_K_GEMM_WO  7
#define L2B_K_GEMM_W13 8
#define L2B_K_GEMM_W2  9
#define L2B_K_GEMM_CLS 10
#define L2B_K_BATCH_EPI 11 /* fused elementwise kernels between the GEMMs (RoPE+KV write,
                              residual+rmsnorm, SwiGLU, logits+argmax)                    */
#define L2B_K_COUNT 12

/* One batch-1 decode step run WITHOUT graph/PDL overlap and with a CUDA event
 * between every two launches: ms_per_class[k] / launches_per_class[k]
 * (L2B_K_COUNT entries each) give the average device time of one launch of
 * kernel class k.  Measurement aid for bench.py's roofline; same results and
 * state effects as l2b_forward.                                                */
int l2b_profile_step(l2b_ctx* ctx, int32_t token, int32_t pos, float* ms_per_class,
                     int32_t* launches_per_class);
/* Same for one step of B sequences (l2b_forward_batch semantics).                        */
int l2b_profile_batch(l2b_ctx* ctx, int32_t B, const int32_t* tokens, const int32_t* pos,
                      float* ms_per_class, int32_t* launches_per_class);

/* Debug/parity taps: copy device RunState buffers (llama2.ts:131-163) of
 * sequence `seq` to host.  Key/value rows come back in the reference's row
 * order (dim floats) whatever the device layout is.                            */
#define L2B_S_X         0   /* x   (dim): residual stream BEFORE the final rmsnorm */
#define L2B_S_KEY_ROW   1   /* key_cache row (layer,pos)   (dim)                   */
#define L2B_S_VALUE_ROW 2   /* value_cache row (layer,pos) (dim)                   */
#define L2B_S_Q         3   /* q   (dim), last layer                               */
#define L2B_S_XB        4   /* xb  (dim): attention output of the last layer       */
#define L2B_S_HB        5   /* hb  (hidden_dim): SwiGLU output of the last layer   */
#define L2B_S_LOGITS    6   /* logits (vocab)                                      */
int l2b_read_state(l2b_ctx* ctx, int32_t which, int32_t seq, int32_t layer, int32_t pos,
                   float* out, uint64_t n_floats);

/* Debug aid: after l2b_set_option("gemm_timeline", 1) the w1/w3 GEMM records clock64()
 * stamps of CTA 0's warp roles per k-block (8 int64 per k-block: producer issue, splitter
 * slot free / tile landed / stored, MMA operands ready / issued); copies up to n of them.  */
int l2b_debug_timeline(l2b_ctx* ctx, int64_t* out, uint64_t n);

/* Forget every sequence: zero the KV cache (newRunState, llama2.ts:160-161) and
 * allow pos to start again from 0.  Weights stay.                              */
int l2b_reset(l2b_ctx* ctx);

/* Tuning knobs (integers; unknown keys return L2B_EINVAL).  See DESIGN.md.      */
int l2b_set_option(l2b_ctx* ctx, const char* key, int64_t value);

/* ---- tensor-parallel wiring (one process per GPU) ---------------------------
 * Each rank exports an opaque handle blob for its exchange buffers
 * (l2b_tp_export: writes up to cap bytes, returns the size), the host exchanges
 * the blobs (torch.distributed / any side channel) and hands all of them, in
 * rank order, to l2b_tp_connect.                                                */
int64_t l2b_tp_export(l2b_ctx* ctx, void* blob, uint64_t cap);
int     l2b_tp_connect(l2b_ctx* ctx, const void* blobs, uint64_t blob_bytes, int32_t n_ranks);

/* ---- native tokenizer (host code, no device) -----------
 * tokenizer.bin parse (llama2.ts:441-449), bpe_encode (llama2.ts:305-344: first-match
 * indexOf semantics, best-score pair merging) and piece lookup (llama2.ts:501-503) for a
 * node-free host.  l2b_tok_encode returns the token count or L2B_EINVAL (unknown character,
 * like the reference's throw at :310, or cap too small).                                  */
typedef struct l2b_tokenizer l2b_tokenizer;
int l2b_tok_load(const uint8_t* data, uint64_t nbytes, int32_t vocab_size, l2b_tokenizer** out);
int l2b_tok_encode(const l2b_tokenizer* tok, const char* text_utf8, int32_t* tokens_out, int32_t cap);
const char* l2b_tok_piece(const l2b_tokenizer* tok, int32_t id);      /* may contain NUL bytes */
int32_t l2b_tok_piece_len(const l2b_tokenizer* tok, int32_t id);
float l2b_tok_score(const l2b_tokenizer* tok, int32_t id);
void l2b_tok_free(l2b_tokenizer* tok);

const char* l2b_last_error(const l2b_ctx* ctx); /* ctx may be NULL: last create error */
int  l2b_abi_version(void);
void l2b_destroy(l2b_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* LLAMA2_B200_H */
