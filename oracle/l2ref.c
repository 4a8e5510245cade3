/*
 * oracle/l2ref.c -- CPU restatement of llama2.ts's transformer() forward pass.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under llama2.ts_b200/ (the product) may
 * import, link or call this file; it is the checker used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY PIN: the reference (/root/reference/llama2.ts) ships no tests, golden
 * vectors or fixtures for this path, and no JavaScript
 * engine exists in the build image, so the reference itself cannot be executed
 * natively.  This restatement is pinned instead against the reference's own
 * SOURCE TEXT executed through oracle/ts_exec.py (a line-level TypeScript ->
 * Python transpiler that reads /root/reference/llama2.ts and runs its
 * functions with JS number semantics); the resulting vectors are committed in
 * tests/golden/ together with the generating script.  Math.exp is libm's exp()
 * in both, not V8's; see DESIGN.md "Oracle".
 *
 * Numerics contract (this IS the definition of "reference result"):
 *   - every JS `number` temporary is an IEEE double;
 *   - every Float32Array store rounds that double to float (RNE);
 *   - loop orders are exactly the reference's;
 *   - compile with -ffp-contract=off so no FMA contraction changes rounding.
 *
 * Each function cites the llama2.ts line range it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__GNUC__)
#define L2REF_API __attribute__((visibility("default")))
#else
#define L2REF_API
#endif

/* ------------------------------------------------------------------------- */
/* primitives: llama2.ts:165-203                                             */

/* llama2.ts:168-170  accum(a,b,size): a[i] += b[i] (f64 add, f32 store) */
L2REF_API void l2ref_accum(float* a, const float* b, int size) {
  for (int i = 0; i < size; i++) a[i] = (float)((double)a[i] + (double)b[i]);
}

/* llama2.ts:172-179  rmsnorm(o,x,weight,size) */
L2REF_API void l2ref_rmsnorm(float* o, const float* x, const float* weight, int size) {
  double ss = 0.0;
  for (int j = 0; j < size; j++) ss += (double)x[j] * (double)x[j];
  ss /= (double)size;
  ss = 1.0 / sqrt(1e-5 + ss);
  for (int j = 0; j < size; j++) o[j] = (float)((double)weight[j] * (ss * (double)x[j]));
}

/* llama2.ts:181-194  softmax(x,xPtr,size), in place */
L2REF_API void l2ref_softmax(float* x, int size) {
  double max_val = (double)x[0];
  for (int i = 1; i < size; i++)
    if ((double)x[i] > max_val) max_val = (double)x[i];
  for (int i = 0; i < size; i++) x[i] = (float)exp((double)x[i] - max_val);
  double sum = 0.0;
  for (int i = 0; i < size; i++) sum += (double)x[i];
  for (int i = 0; i < size; i++) x[i] = (float)((double)x[i] / sum);
}

/* llama2.ts:196-203  matmul(xout,x,w,n,d): W(d,n) @ x(n) -> xout(d) */
/* Rows are independent and each row keeps the reference's sequential j order,
 * so splitting ROWS over threads cannot change a single bit of the result.
 * l2ref_set_threads(1) (the default) is the reference's single JS thread. */
static int g_threads = 1;
L2REF_API void l2ref_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
L2REF_API int l2ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

L2REF_API void l2ref_matmul(float* xout, const float* x, const float* w, int n, int d) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)
#endif
  for (int i = 0; i < d; i++) {
    double sum = 0.0;
    const float* wi = w + (size_t)i * (size_t)n;
    for (int j = 0; j < n; j++) sum += (double)wi[j] * (double)x[j];
    xout[i] = (float)sum;
  }
}

/* ------------------------------------------------------------------------- */
/* model: Config (llama2.ts:69-93), TransformerWeights (:95-129),            */
/* RunState (:131-163)                                                       */

typedef struct l2ref_model {
  int dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len;
  int shared_weights, head_size;
  /* weights: borrowed pointers, laid out as readWeights() slices them */
  const float* token_embedding_table; /* (vocab, dim) */
  const float* rms_att_weight;        /* (layer, dim) */
  const float* wq;                    /* (layer, dim, dim) */
  const float* wk;
  const float* wv;
  const float* wo;
  const float* rms_ffn_weight;        /* (layer, dim) */
  const float* w1;                    /* (layer, hidden, dim) */
  const float* w2;                    /* (layer, dim, hidden) */
  const float* w3;                    /* (layer, hidden, dim) */
  const float* rms_final_weight;      /* (dim) */
  const float* freq_cis_real;         /* (seq_len, head_size/2) */
  const float* freq_cis_imag;
  const float* wcls;                  /* (vocab, dim) */
  /* run state, owned; llama2.ts:147-163 (all zero-initialised) */
  float *x, *xb, *xb2, *hb, *hb2, *q, *k, *v, *att, *logits;
  float *key_cache, *value_cache;
  void* blob; /* non-NULL when the model owns a copy of the checkpoint */
} l2ref_model;

/* llama2.ts:80-93 readConfig + :147-163 newRunState.  hdr = the 7 header ints. */
L2REF_API l2ref_model* l2ref_create(const int32_t hdr[7]) {
  l2ref_model* m = (l2ref_model*)calloc(1, sizeof(l2ref_model));
  if (!m) return NULL;
  m->dim = hdr[0];
  m->hidden_dim = hdr[1];
  m->n_layers = hdr[2];
  m->n_heads = hdr[3];
  m->n_kv_heads = hdr[4]; /* read, then ignored, exactly like the reference */
  m->vocab_size = hdr[5] < 0 ? -hdr[5] : hdr[5];
  m->seq_len = hdr[6];
  m->shared_weights = hdr[5] > 0;
  m->head_size = m->dim / m->n_heads;
  size_t D = (size_t)m->dim, F = (size_t)m->hidden_dim, V = (size_t)m->vocab_size;
  size_t kv = (size_t)m->n_layers * (size_t)m->seq_len * D;
  m->x = (float*)calloc(D, 4);
  m->xb = (float*)calloc(D, 4);
  m->xb2 = (float*)calloc(D, 4);
  m->hb = (float*)calloc(F, 4);
  m->hb2 = (float*)calloc(F, 4);
  m->q = (float*)calloc(D, 4);
  m->k = (float*)calloc(D, 4);
  m->v = (float*)calloc(D, 4);
  m->att = (float*)calloc((size_t)m->n_heads * (size_t)m->seq_len, 4);
  m->logits = (float*)calloc(V, 4);
  m->key_cache = (float*)calloc(kv, 4);
  m->value_cache = (float*)calloc(kv, 4);
  return m;
}

/* Total number of floats that follow the 28-byte header (llama2.ts:112-129). */
L2REF_API uint64_t l2ref_weight_floats(const l2ref_model* m) {
  uint64_t D = m->dim, F = m->hidden_dim, L = m->n_layers, V = m->vocab_size, S = m->seq_len;
  uint64_t hs2 = (uint64_t)(m->head_size / 2);
  uint64_t n = V * D + L * D + 4 * L * D * D + L * D + 3 * L * F * D + D + 2 * S * hs2;
  if (!m->shared_weights) n += V * D;
  return n;
}

/* llama2.ts:112-129 readWeights: slice one contiguous float blob in file order. */
L2REF_API int l2ref_bind_weights(l2ref_model* m, const float* p) {
  size_t D = m->dim, F = m->hidden_dim, L = m->n_layers, V = m->vocab_size, S = m->seq_len;
  size_t hs2 = (size_t)(m->head_size / 2);
  m->token_embedding_table = p; p += V * D;
  m->rms_att_weight = p;        p += L * D;
  m->wq = p;                    p += L * D * D;
  m->wk = p;                    p += L * D * D;
  m->wv = p;                    p += L * D * D;
  m->wo = p;                    p += L * D * D;
  m->rms_ffn_weight = p;        p += L * D;
  m->w1 = p;                    p += L * F * D;
  m->w2 = p;                    p += L * D * F;
  m->w3 = p;                    p += L * F * D;
  m->rms_final_weight = p;      p += D;
  m->freq_cis_real = p;         p += S * hs2;
  m->freq_cis_imag = p;         p += S * hs2;
  m->wcls = m->shared_weights ? m->token_embedding_table : p; /* llama2.ts:127 */
  return 0;
}

L2REF_API void l2ref_destroy(l2ref_model* m) {
  if (!m) return;
  free(m->x); free(m->xb); free(m->xb2); free(m->hb); free(m->hb2);
  free(m->q); free(m->k); free(m->v); free(m->att); free(m->logits);
  free(m->key_cache); free(m->value_cache);
  free(m->blob);
  free(m);
}

L2REF_API float* l2ref_logits(l2ref_model* m) { return m->logits; }
L2REF_API float* l2ref_x(l2ref_model* m) { return m->x; }
L2REF_API float* l2ref_key_cache(l2ref_model* m) { return m->key_cache; }
L2REF_API float* l2ref_value_cache(l2ref_model* m) { return m->value_cache; }
L2REF_API int l2ref_vocab(const l2ref_model* m) { return m->vocab_size; }

/* llama2.ts:224-235  RoPE on q and k for one position (adjacent pairs). */
L2REF_API void l2ref_rope(float* q, float* k, const float* fcr_tab, const float* fci_tab,
                          int pos, int dim, int head_size) {
  for (int i = 0; i < dim; i += 2) {
    double q0 = q[i], q1 = q[i + 1], k0 = k[i], k1 = k[i + 1];
    double fcr = fcr_tab[pos * head_size / 2 + (i % head_size) / 2];
    double fci = fci_tab[pos * head_size / 2 + (i % head_size) / 2];
    q[i] = (float)(q0 * fcr - q1 * fci);
    q[i + 1] = (float)(q0 * fci + q1 * fcr);
    k[i] = (float)(k0 * fcr - k1 * fci);
    k[i + 1] = (float)(k0 * fci + k1 * fcr);
  }
}

/* llama2.ts:244-267  multi-head attention for one layer at position pos. */
static void attention(l2ref_model* m, size_t loff, int pos) {
  const int dim = m->dim, hs = m->head_size, S = m->seq_len;
  for (int h = 0; h < m->n_heads; h++) {
    const float* q = m->q + h * hs;
    float* att = m->att + (size_t)h * S;
    for (int t = 0; t <= pos; t++) {
      const float* ck = m->key_cache + loff + (size_t)t * dim + h * hs;
      double score = 0.0;
      for (int i = 0; i < hs; i++) score += (double)q[i] * (double)ck[i];
      att[t] = (float)(score / sqrt((double)hs));
    }
    l2ref_softmax(att, pos + 1);
    for (int i = 0; i < hs; i++) m->xb[h * hs + i] = 0.0f;
    for (int t = 0; t <= pos; t++) {
      const double att_t = att[t];
      const float* cv = m->value_cache + loff + (size_t)t * dim + h * hs;
      for (int i = 0; i < hs; i++)
        m->xb[h * hs + i] = (float)((double)m->xb[h * hs + i] + att_t * (double)cv[i]);
    }
  }
}

/* llama2.ts:205-303  transformer(token,pos,p,s,w) */
L2REF_API int l2ref_forward(l2ref_model* m, int token, int pos) {
  const int dim = m->dim, hidden = m->hidden_dim;
  if (token < 0 || token >= m->vocab_size || pos < 0 || pos >= m->seq_len) return -1;
  float* x = m->x;
  memcpy(x, m->token_embedding_table + (size_t)token * dim, (size_t)dim * 4); /* :211 */
  for (int l = 0; l < m->n_layers; l++) {
    l2ref_rmsnorm(m->xb, x, m->rms_att_weight + (size_t)l * dim, dim);        /* :216 */
    l2ref_matmul(m->q, m->xb, m->wq + (size_t)l * dim * dim, dim, dim);       /* :219 */
    l2ref_matmul(m->k, m->xb, m->wk + (size_t)l * dim * dim, dim, dim);
    l2ref_matmul(m->v, m->xb, m->wv + (size_t)l * dim * dim, dim, dim);
    l2ref_rope(m->q, m->k, m->freq_cis_real, m->freq_cis_imag, pos, dim, m->head_size);
    size_t loff = (size_t)l * m->seq_len * dim;                               /* :238 */
    memcpy(m->key_cache + loff + (size_t)pos * dim, m->k, (size_t)dim * 4);
    memcpy(m->value_cache + loff + (size_t)pos * dim, m->v, (size_t)dim * 4);
    attention(m, loff, pos);                                                  /* :244 */
    l2ref_matmul(m->xb2, m->xb, m->wo + (size_t)l * dim * dim, dim, dim);     /* :270 */
    l2ref_accum(x, m->xb2, dim);                                              /* :273 */
    l2ref_rmsnorm(m->xb, x, m->rms_ffn_weight + (size_t)l * dim, dim);        /* :276 */
    l2ref_matmul(m->hb, m->xb, m->w1 + (size_t)l * hidden * dim, dim, hidden);
    l2ref_matmul(m->hb2, m->xb, m->w3 + (size_t)l * hidden * dim, dim, hidden);
    for (int i = 0; i < hidden; i++) {                                        /* :284 */
      double h = m->hb[i];
      m->hb[i] = (float)(h * (1.0 / (1.0 + exp(-h))));
    }
    for (int i = 0; i < hidden; i++)                                          /* :289 */
      m->hb[i] = (float)((double)m->hb[i] * (double)m->hb2[i]);
    l2ref_matmul(m->xb, m->hb, m->w2 + (size_t)l * dim * hidden, hidden, dim);/* :292 */
    l2ref_accum(x, m->xb, dim);                                               /* :295 */
  }
  l2ref_rmsnorm(x, x, m->rms_final_weight, dim);                              /* :299 */
  l2ref_matmul(m->logits, x, m->wcls, dim, m->vocab_size);                    /* :302 */
  return 0;
}

/* ------------------------------------------------------------------------- */
/* host-side pieces of the reference needed to drive the path without node   */

/* llama2.ts:348-354  xorshift64* on a 64-bit seed */
L2REF_API uint32_t l2ref_random_u32(uint64_t* seed) {
  uint64_t s = *seed;
  s ^= s >> 12;
  s ^= s << 25;
  s ^= s >> 27;
  *seed = s;
  return (uint32_t)((s * 0x2545F4914F6CDD1DULL) >> 32);
}

/* llama2.ts:356-360  (u32 / 256) / 16777216.0 in f64, forced to f32 */
L2REF_API float l2ref_random_f32(uint64_t* seed) {
  return (float)(((double)l2ref_random_u32(seed) / 256.0) / 16777216.0);
}

/* llama2.ts:364-366  argmax: first maximum wins (strict >, NaN never wins) */
L2REF_API int l2ref_argmax(const float* arr, int n) {
  int max_idx = 0;
  for (int i = 0; i < n; i++)
    if (arr[i] > arr[max_idx]) max_idx = i;
  return max_idx;
}

/* llama2.ts:368-376  sample(): CDF walk scaled by the f64 sum */
L2REF_API int l2ref_sample(const float* probs, int n, uint64_t* seed) {
  double sum = 0.0;
  for (int i = 0; i < n; i++) sum += (double)probs[i];
  double r = (double)l2ref_random_f32(seed) * sum;
  double cum = 0.0;
  for (int i = 0; i < n; i++) {
    cum += (double)probs[i];
    if (r < cum) return i;
  }
  return 0;
}

typedef struct { float prob; int index; } probindex_t;

/* Array.prototype.sort is stable (ES2019); comparator is b.prob - a.prob. */
static void stable_sort_desc(probindex_t* a, probindex_t* tmp, int n) {
  if (n < 2) return;
  int mid = n / 2;
  stable_sort_desc(a, tmp, mid);
  stable_sort_desc(a + mid, tmp, n - mid);
  int i = 0, j = mid, k = 0;
  while (i < mid && j < n) {
    /* take right only when strictly greater: keeps equal elements in order */
    if (a[j].prob > a[i].prob) tmp[k++] = a[j++]; else tmp[k++] = a[i++];
  }
  while (i < mid) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, (size_t)n * sizeof(probindex_t));
}

/* llama2.ts:378-394  sample_topp(): note the exclusive `i < lastIdx` walk and
 * the fallback to token 0 -- quirks reproduced on purpose. */
L2REF_API int l2ref_sample_topp(const float* probs, int n, double topp, uint64_t* seed) {
  probindex_t* pi = (probindex_t*)malloc((size_t)n * sizeof(probindex_t) * 2);
  for (int i = 0; i < n; i++) { pi[i].index = i; pi[i].prob = probs[i]; }
  stable_sort_desc(pi, pi + n, n);
  double cum = 0.0;
  int last = 0;
  for (int i = 0; i < n; i++) {
    cum += (double)pi[i].prob;
    if (cum > topp) { last = i; break; }
  }
  double r = (double)l2ref_random_f32(seed) * cum;
  cum = 0.0;
  int ret = 0;
  for (int i = 0; i < last; i++) {
    cum += (double)pi[i].prob;
    if (r < cum) { ret = pi[i].index; break; }
  }
  free(pi);
  return ret;
}

/* llama2.ts:476-494  temperature + softmax + sampler choice on logits (in place) */
L2REF_API int l2ref_sample_next(float* logits, int n, double temperature, double topp,
                                uint64_t* seed) {
  if (temperature == 0.0) return l2ref_argmax(logits, n);
  for (int q = 0; q < n; q++) logits[q] = (float)((double)logits[q] / temperature);
  l2ref_softmax(logits, n);
  if (topp <= 0 || topp >= 1) return l2ref_sample(logits, n, seed);
  return l2ref_sample_topp(logits, n, topp, seed);
}

/* llama2.ts:460-508  the generate loop (no printing).  prompt = forced tokens
 * (may be NULL), out_tokens receives `next` for every executed step, and when
 * logits_out != NULL the V logits of every step are appended to it.  Returns the
 * number of steps executed (the loop breaks after next == 1, llama2.ts:499). */
L2REF_API int l2ref_generate(l2ref_model* m, int steps, const int32_t* prompt, int n_prompt,
                             double temperature, double topp, uint64_t seed,
                             int32_t* out_tokens, float* logits_out) {
  if (steps <= 0 || steps > m->seq_len) steps = m->seq_len; /* llama2.ts:439 */
  int token = 1, pos = 0, next = 0, n = 0;
  while (pos < steps) {
    if (l2ref_forward(m, token, pos) != 0) return -1;
    if (logits_out) memcpy(logits_out + (size_t)n * m->vocab_size, m->logits, (size_t)m->vocab_size * 4);
    if (pos < n_prompt) next = prompt[pos];
    else next = l2ref_sample_next(m->logits, m->vocab_size, temperature, topp, &seed);
    out_tokens[n++] = next;
    pos++;
    if (next == 1) break;
    token = next;
  }
  return n;
}

/* ------------------------------------------------------------------------- */
/* tokenizer.bin + bpe_encode: llama2.ts:441-449, :305-344                   */

typedef struct l2ref_tokenizer {
  int vocab_size;
  char** vocab;     /* NUL-terminated copies of the raw token bytes */
  int* lens;
  float* scores;
} l2ref_tokenizer;

L2REF_API l2ref_tokenizer* l2ref_tokenizer_load(const uint8_t* buf, uint64_t nbytes, int vocab_size) {
  l2ref_tokenizer* t = (l2ref_tokenizer*)calloc(1, sizeof(*t));
  t->vocab_size = vocab_size;
  t->vocab = (char**)calloc((size_t)vocab_size, sizeof(char*));
  t->lens = (int*)calloc((size_t)vocab_size, sizeof(int));
  t->scores = (float*)calloc((size_t)vocab_size, sizeof(float));
  uint64_t p = 4; /* int32 max_token_length, ignored (llama2.ts:445) */
  for (int i = 0; i < vocab_size; i++) {
    if (p + 8 > nbytes) { t->vocab_size = i; break; }
    int32_t len;
    memcpy(&t->scores[i], buf + p, 4); p += 4;
    memcpy(&len, buf + p, 4); p += 4;
    t->vocab[i] = (char*)calloc((size_t)len + 1, 1);
    memcpy(t->vocab[i], buf + p, (size_t)len); p += (uint64_t)len;
    t->lens[i] = len;
  }
  return t;
}

L2REF_API void l2ref_tokenizer_free(l2ref_tokenizer* t) {
  if (!t) return;
  for (int i = 0; i < t->vocab_size; i++) free(t->vocab[i]);
  free(t->vocab); free(t->lens); free(t->scores); free(t);
}

L2REF_API const char* l2ref_tokenizer_piece(const l2ref_tokenizer* t, int id) { return t->vocab[id]; }
L2REF_API float l2ref_tokenizer_score(const l2ref_tokenizer* t, int id) { return t->scores[id]; }

/* vocab.indexOf(str): first match wins (llama2.ts:309,323) */
static int vocab_index_of(const l2ref_tokenizer* t, const char* s, int len) {
  for (int i = 0; i < t->vocab_size; i++)
    if (t->lens[i] == len && memcmp(t->vocab[i], s, (size_t)len) == 0) return i;
  return -1;
}

/* llama2.ts:305-344 bpe_encode.  ASCII prompts only: the reference iterates
 * UTF-16 code units (text.charAt), which coincide with bytes for ASCII.
 * Returns n_tokens, or -1 when a character is not in the vocab (:310 throws). */
L2REF_API int l2ref_bpe_encode(const l2ref_tokenizer* t, const char* text, int32_t* tokens) {
  int n = 0;
  for (const char* c = text; *c; ++c) {
    int id = vocab_index_of(t, c, 1);
    if (id == -1) return -1;
    tokens[n++] = id;
  }
  char* buf = (char*)malloc(1024);
  for (;;) {
    double best_score = -1e10;
    int best_id = -1, best_idx = -1;
    for (int i = 0; i < n - 1; ++i) {
      int la = t->lens[tokens[i]], lb = t->lens[tokens[i + 1]];
      if (la + lb >= 1024) continue;
      memcpy(buf, t->vocab[tokens[i]], (size_t)la);
      memcpy(buf + la, t->vocab[tokens[i + 1]], (size_t)lb);
      int id = vocab_index_of(t, buf, la + lb);
      if (id != -1 && (double)t->scores[id] > best_score) {
        best_score = t->scores[id];
        best_id = id;
        best_idx = i;
      }
    }
    if (best_idx == -1) break;
    tokens[best_idx] = best_id;
    for (int i = best_idx + 1; i < n - 1; i++) tokens[i] = tokens[i + 1];
    n--;
  }
  free(buf);
  return n;
}
