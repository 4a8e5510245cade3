// stream_kernel.cuh -- EXPERIMENT (option mega=2, off by default): batch-1 decode
// (llama2.ts:205-303) and the greedy loop around it (llama2.ts:465-508) as ONE persistent
// kernel without grid barriers.
//
// Idea: the per-op kernels lose a few microseconds per kernel boundary because the HBM
// stream stops while activations change hands.  Here
//  * weights: every warp owns a private ring of shared-memory stages that lane 0 fills with
//    1-D bulk async copies (cp.async.bulk, completion on an mbarrier).  The ring follows the
//    warp's weight stream ACROSS phase, layer and token boundaries -- weights never depend
//    on activations -- so the copies for the next phase are in flight during the hand-over.
//  * activations: each phase publishes its outputs to L2 as 8-byte {value, sequence} words
//    ("LL", as in the tensor-parallel exchange); consumers poll the words they need.  x, hb,
//    attention partials and argmax candidates go to one mailbox per consumer CTA.
//
// Outcome (B200): parity-exact like the default path, but slower --
// stories15M 323 us/token against 120 us, Llama-2-7B 11.4 ms against 4.7 ms.  The
// microbenchmark tools/llbench.cu shows why: one all-to-all hand-over between 148 CTAs
// costs 2-4 us whether it is done with LL words or with a grid barrier, the same as a whole
// kernel boundary under programmatic dependent launch (3.7 us), and a persistent kernel adds
// the serial latency of a lone warp per phase (fp64 sqrt/divide/exp, reductions) that the
// per-op kernels overlap across their prologues.  Kept as a tested option and as evidence.
//
// Phases per layer: rmsnorm + q/k/v rows + RoPE + KV write | attention (split over heads
// and, for long contexts, over time chunks; partial results merged by the consumers) |
// wo + residual | rmsnorm + w1/w3 + SwiGLU | w2 + residual; then the classifier, argmax
// and state advance.  GEMV arithmetic and summation order are those of decode_kernels.cuh.
#pragma once
#include <math.h>

#include "../llama2.ts_b200/csrc/common.cuh"
#include "../llama2.ts_b200/csrc/decode_kernels.cuh"
#include "mega_kernel.cuh"

namespace l2b {

constexpr int kSThreads = 512;
constexpr int kSWarps = kSThreads / 32;
constexpr int kSSeg = 512;                  // floats of one row in a ring stage
constexpr int kSStageFloats = 2 * kSSeg;    // a stage: two row segments, or whole small row pairs
constexpr int kSMaxStages = 6;
constexpr int kSMaxChunks = 4;              // attention: time chunks per head
constexpr int kSMaxOwn = 256;               // residual rows one CTA owns in the wo / w2 phases
constexpr int kSMaxV4 = 6;                  // float4 per thread the hidden-vector gather keeps in registers
constexpr int kSMaxV4D = 4;                 // same for vectors of length dim

enum { SK_QKV = 0, SK_ATT = 1, SK_XWO = 2, SK_HB = 3, SK_XW2 = 4, SK_KINDS = 8 };

struct StreamParams {
  int D, F, L, H, hs, V, steps;
  const float *tok_emb, *rms_att, *wqkv, *wo, *rms_ffn, *w13, *w2, *rms_final, *fcr, *fci, *wcls;
  float *kc, *vc;          // sequence 0 of the head-major cache
  long long kv_layer;      // floats per layer of one cache
  float* logits;
  float* x_out;            // residual stream of the last step (state read-back)
  int* ctl;
  int* next;
  const int* forced;
  int* out_tokens;
  // LL words.  q/k/v have few readers (the CTAs of one head) and live in one shared copy;
  // x, hb, the attention partials and the argmax candidates are read by every CTA: the
  // producer stores them into one mailbox per consumer CTA (mbox_stride words apart), so a
  // word is only ever polled by the threads of one CTA -- no hot lines in L2.
  uint2 *q_ll, *k_ll, *v_ll;
  uint2 *x_ll, *hb_ll, *part_ll, *am_ll;   // offsets inside mailbox 0
  long long mbox_stride;
  int part_stride;         // words per attention partial (hs + 4)
  int* err;
  unsigned seq0;
  int n_steps;
  int evict_first;
  int stages;              // ring stages per warp
  int gmax;                // attention chunks per head (<= kSMaxChunks)
  long long* dbg;          // option gemv_timeline: [event][cta 0 | last cta][6] globaltimer stamps of the last step
};

__device__ __forceinline__ void st_gpu_u4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void st_gpu_u2(void* p, uint32_t a, uint32_t b) {
  asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ int ld_volatile_i32(const int* p) {
  int r;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                              uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

// LL polling.  Bounded: a lost peer sets *err and every poller gives up (the host reports
// L2B_ECOMM), the kernel always terminates.
__device__ __forceinline__ bool ll_ok(const uint4& v, uint32_t seq) { return v.y == seq && v.w == seq; }
__device__ __forceinline__ bool ll_give_up(unsigned& spins, int* err) {
  if ((spins++ & 4095u) == 0u) {
    if (ld_volatile_i32(err) != 0) return true;
    if (spins > (1u << 24)) {
      atomicExch(err, 1);
      return true;
    }
  }
  return false;
}
// N 16-byte units (two LL words each) at ptr[0..N): all loads are issued before the first is
// checked, invalid ones are re-issued together until every unit carries `seq`.
template <int N>
__device__ __forceinline__ void ll_wait_units(const uint4* const (&ptr)[N], const bool (&on)[N], uint4 (&v)[N],
                                              uint32_t seq, int* err) {
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (on[i]) v[i] = ld_volatile_u4(ptr[i]);
  unsigned spins = 0;
  while (true) {
    bool pending = false;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (on[i] && !ll_ok(v[i], seq)) {
        v[i] = ld_volatile_u4(ptr[i]);
        pending = true;
      }
    }
    if (!pending) return;
    if (ll_give_up(spins, err)) return;
  }
}
__device__ __forceinline__ uint4 ll_wait2(const uint2* words, uint32_t seq, int* err) {
  const uint4* const ptr[1] = {reinterpret_cast<const uint4*>(words)};
  const bool on[1] = {true};
  uint4 v[1];
  ll_wait_units<1>(ptr, on, v, seq, err);
  return v[0];
}
__device__ __forceinline__ float4 ll_f4(const uint4& a, const uint4& b) {
  return make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
}
// n floats (n4 float4, thread-strided) of an LL vector into registers
template <int K>
__device__ __forceinline__ void ll_gather_vec(const uint2* src_ll, int n4, uint32_t seq, int* err, float4 (&v)[K]) {
  const uint4* b = reinterpret_cast<const uint4*>(src_ll) + 2 * (size_t)threadIdx.x;
  uint4 u[2 * K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    if ((int)threadIdx.x + k * kSThreads < n4) {
      u[2 * k] = ld_volatile_u4(b + 2 * k * kSThreads);
      u[2 * k + 1] = ld_volatile_u4(b + 2 * k * kSThreads + 1);
    }
  }
  unsigned spins = 0;
  while (true) {
    bool pending = false;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if ((int)threadIdx.x + k * kSThreads < n4) {
        if (!ll_ok(u[2 * k], seq)) {
          u[2 * k] = ld_volatile_u4(b + 2 * k * kSThreads);
          pending = true;
        }
        if (!ll_ok(u[2 * k + 1], seq)) {
          u[2 * k + 1] = ld_volatile_u4(b + 2 * k * kSThreads + 1);
          pending = true;
        }
      }
    }
    if (!pending) break;
    if (ll_give_up(spins, err)) break;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = ll_f4(u[2 * k], u[2 * k + 1]);
}

// The five weight streams of a step.  Everything the hot path needs about them (row ranges
// per CTA and per warp, stage packing) is tabulated once at kernel start: a lone warp running
// integer divisions between two stages costs microseconds here.
enum { SP_QKV = 0, SP_WO = 1, SP_W13 = 2, SP_W2 = 3, SP_CLS = 4, SP_KINDS = 5 };
struct SShared {
  double red[kSWarps];
  float wmax[kSWarps];
  double wsum[kSWarps];
  float bv[kSWarps];
  int bi[kSWarps];
  float xown[kSMaxOwn];
  int tok, pos;
  int rng[SP_KINDS][kSWarps][2];  // row pairs [a, b) of warp w in a phase of this kind
  int cta[SP_KINDS][2];           // same for the whole CTA
  int per[SP_KINDS];              // whole row pairs per ring stage (rows of <= kSSeg floats), else 0
  int n[SP_KINDS];                // row length
};
__device__ __forceinline__ int s_kind(const StreamParams& p, int idx) { return idx == 4 * p.L ? SP_CLS : (idx & 3); }
__device__ __forceinline__ const float* s_weights(const StreamParams& p, int idx) {
  const size_t D = p.D, F = p.F;
  if (idx == 4 * p.L) return p.wcls;
  const size_t l = (size_t)(idx >> 2);
  switch (idx & 3) {
    case 0: return p.wqkv + l * 3 * D * D;
    case 1: return p.wo + l * D * D;
    case 2: return p.w13 + l * 2 * F * D;
    default: return p.w2 + l * D * F;
  }
}
__device__ void s_tabulate(const StreamParams& p, SShared& sh) {
  const int t = threadIdx.x;
  if (t < SP_KINDS * kSWarps) {
    const int kind = t / kSWarps, warp = t % kSWarps;
    const int rows = kind == SP_QKV ? 3 * p.D : kind == SP_W13 ? 2 * p.F : kind == SP_CLS ? p.V : p.D;
    const int n = kind == SP_W2 ? p.F : p.D;
    const int npairs = rows >> 1;
    const int c0 = (int)(((long long)npairs * blockIdx.x) / gridDim.x);
    const int c1 = (int)(((long long)npairs * (blockIdx.x + 1)) / gridDim.x);
    const int cnt = c1 - c0;
    sh.rng[kind][warp][0] = c0 + (cnt * warp) / kSWarps;
    sh.rng[kind][warp][1] = c0 + (cnt * (warp + 1)) / kSWarps;
    if (warp == 0) {
      sh.cta[kind][0] = c0;
      sh.cta[kind][1] = c1;
      sh.per[kind] = n <= kSSeg ? kSStageFloats / (2 * n) : 0;
      sh.n[kind] = n;
    }
  }
}

// Per-warp state: the fetch cursor (what lane 0 requests next) and the ring position the
// consumer side is at.  Uniform across the warp.
struct SWarp {
  // fetch cursor
  int f_ph, f_step, f_pair, f_end, f_seg, f_n, f_per;
  const float* f_W;
  int f_slot;
  bool f_done;
  // consumer
  int c_slot;
  uint32_t c_parity;   // bit s: parity to wait for on stage s
  float* ring;         // this warp's stages
  uint64_t* bars;
  int stages;
  uint64_t pol;
};

__device__ __forceinline__ void s_cursor_enter(const StreamParams& p, const SShared& sh, SWarp& w) {
  // position the cursor on the first non-empty phase at or after (f_step, f_ph)
  const int warp = threadIdx.x >> 5;
  while (true) {
    if (w.f_step >= p.n_steps) {
      w.f_done = true;
      return;
    }
    const int kind = s_kind(p, w.f_ph);
    const int a = sh.rng[kind][warp][0], b = sh.rng[kind][warp][1];
    if (a < b) {
      w.f_pair = a;
      w.f_end = b;
      w.f_seg = 0;
      w.f_n = sh.n[kind];
      w.f_per = sh.per[kind];
      w.f_W = s_weights(p, w.f_ph);
      return;
    }
    if (++w.f_ph == 4 * p.L + 1) {
      w.f_ph = 0;
      ++w.f_step;
    }
  }
}
// request the next stage of this warp's weight stream into ring slot f_slot
__device__ __noinline__ void s_fetch(const StreamParams& p, const SShared& sh, SWarp& w) {
  if (w.f_done) return;
  const int lane = threadIdx.x & 31;
  float* dst = w.ring + (size_t)w.f_slot * kSStageFloats;
  uint64_t* bar = w.bars + w.f_slot;
  const int n = w.f_n;
  if (w.f_per > 0) {
    const int left = w.f_end - w.f_pair;
    const int np = w.f_per < left ? w.f_per : left;
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)np * 2u * (uint32_t)n * 4u;
      mbar_arrive_expect_tx(bar, bytes);
      bulk_g2s_hint(dst, w.f_W + (size_t)w.f_pair * 2 * n, bytes, bar, w.pol);
    }
    w.f_pair += np;
  } else {
    const int off = w.f_seg * kSSeg;
    const int len = (n - off) < kSSeg ? (n - off) : kSSeg;
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)len * 4u;
      mbar_arrive_expect_tx(bar, 2u * bytes);
      const float* r0 = w.f_W + (size_t)(2 * w.f_pair) * n + off;
      bulk_g2s_hint(dst, r0, bytes, bar, w.pol);
      bulk_g2s_hint(dst + kSSeg, r0 + n, bytes, bar, w.pol);
    }
    if (off + len >= n) {
      w.f_seg = 0;
      ++w.f_pair;
    } else {
      ++w.f_seg;
    }
  }
  if (++w.f_slot == w.stages) w.f_slot = 0;
  if (w.f_pair == w.f_end) {
    if (++w.f_ph == 4 * p.L + 1) {
      w.f_ph = 0;
      ++w.f_step;
    }
    s_cursor_enter(p, sh, w);
  }
}
__device__ __forceinline__ const float4* s_acquire(SWarp& w) {
  mbar_wait(w.bars + w.c_slot, (w.c_parity >> w.c_slot) & 1u);
  return reinterpret_cast<const float4*>(w.ring + (size_t)w.c_slot * kSStageFloats);
}
__device__ __forceinline__ void s_release(const StreamParams& p, const SShared& sh, SWarp& w) {
  w.c_parity ^= 1u << w.c_slot;
  if (++w.c_slot == w.stages) w.c_slot = 0;
  __syncwarp();
  s_fetch(p, sh, w);  // refill the slot just drained
}



// ---- activation gathers (whole CTA) -------------------------------------------------------

// x (LL words, or the embedding row for layer 0) -> rmsnorm -> xs; also refreshes the CTA's
// own residual rows.  llama2.ts:172-179, 211, 216, 274, 297.
__device__ __noinline__ void s_gather_rms(const StreamParams& p, const uint2* src_ll, uint32_t seq,
                                          const float* src_plain, const float* rms_w, unsigned char* xs,
                                          SShared& sh, float* write_x) {
  typedef XVec<true> XV;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = p.D, n4 = n >> 2;
  const int own0 = 2 * sh.cta[SP_WO][0], own1 = 2 * sh.cta[SP_WO][1];
  float4 v[kSMaxV4D];
  if (src_plain) {
#pragma unroll
    for (int k = 0; k < kSMaxV4D; ++k) {
      const int j = threadIdx.x + k * kSThreads;
      if (j < n4) v[k] = __ldg(reinterpret_cast<const float4*>(src_plain) + j);
    }
  } else {
    ll_gather_vec<kSMaxV4D>(src_ll, n4, seq, p.err, v);
  }
  double ss = 0.0;
#pragma unroll
  for (int k = 0; k < kSMaxV4D; ++k) {
    const int j = threadIdx.x + k * kSThreads;
    if (j < n4) {
      ss += (double)v[k].x * (double)v[k].x + (double)v[k].y * (double)v[k].y + (double)v[k].z * (double)v[k].z +
            (double)v[k].w * (double)v[k].w;
      if (write_x != nullptr) reinterpret_cast<float4*>(write_x)[j] = v[k];
    }
  }
  ss = warp_sum_f64(ss);
  __syncthreads();  // every warp is done with the previous contents of xs / red / xown
  if (lane == 0) sh.red[warp] = ss;
#pragma unroll
  for (int k = 0; k < kSMaxV4D; ++k) {
    const int r = 4 * (threadIdx.x + k * kSThreads);
    if (r < n && r + 3 >= own0 && r < own1) {
      const float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (r + c >= own0 && r + c < own1) sh.xown[r + c - own0] = e[c];
    }
  }
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < kSWarps; ++w) tot += sh.red[w];
  tot /= (double)n;
  tot = 1.0 / sqrt(1e-5 + tot);
  const float4* rw4 = reinterpret_cast<const float4*>(rms_w);
#pragma unroll
  for (int k = 0; k < kSMaxV4D; ++k) {
    const int j = threadIdx.x + k * kSThreads;
    if (j < n4) {
      const float4 w = __ldg(rw4 + j);
      float4 o;
      o.x = (float)((double)w.x * (tot * (double)v[k].x));
      o.y = (float)((double)w.y * (tot * (double)v[k].y));
      o.z = (float)((double)w.z * (tot * (double)v[k].z));
      o.w = (float)((double)w.w * (tot * (double)v[k].w));
      XV::store(xs, n4, j, o);
    }
  }
  __syncthreads();
}

// hb (LL words) -> xs.  llama2.ts:289
__device__ __noinline__ void s_gather_copy(const StreamParams& p, const uint2* src_ll, uint32_t seq, int n,
                                           unsigned char* xs) {
  typedef XVec<true> XV;
  const int n4 = n >> 2;
  float4 v[kSMaxV4];
  ll_gather_vec<kSMaxV4>(src_ll, n4, seq, p.err, v);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSMaxV4; ++k) {
    const int j = threadIdx.x + k * kSThreads;
    if (j < n4) XV::store(xs, n4, j, v[k]);
  }
  __syncthreads();
}

// attention partials -> xb -> xs.  Each partial is a locally normalised weighted sum o_g with
// its score maximum m_g and exp-sum s_g; xb = sum_g o_g * (e^(m_g-M) s_g / sum_g' e^(m_g'-M) s_g').
// One chunk: the weight is exactly 1 and xb is the partial itself.
__device__ __noinline__ void s_gather_attn(const StreamParams& p, uint32_t seq, int G, unsigned char* xs) {
  typedef XVec<true> XV;
  const int n4 = p.D >> 2;
  float4 v[kSMaxV4D];
#pragma unroll 1
  for (int k = 0; k < kSMaxV4D; ++k) {
    const int j = threadIdx.x + k * kSThreads;
    if (j < n4) {
      const int h = (4 * j) / p.hs, c = 4 * j - h * p.hs;
      const uint2* base = p.part_ll + (size_t)blockIdx.x * p.mbox_stride + (size_t)h * kSMaxChunks * p.part_stride;
      // per chunk: the four output values (2 units) and {max, sum.lo | sum.hi, pad} (2 units)
      const uint4* ptr[4 * kSMaxChunks];
      bool on[4 * kSMaxChunks];
      uint4 u[4 * kSMaxChunks];
#pragma unroll
      for (int g = 0; g < kSMaxChunks; ++g) {
        const uint2* pg = base + (size_t)g * p.part_stride;
        ptr[4 * g] = reinterpret_cast<const uint4*>(pg + c);
        ptr[4 * g + 1] = reinterpret_cast<const uint4*>(pg + c + 2);
        ptr[4 * g + 2] = reinterpret_cast<const uint4*>(pg + p.hs);
        ptr[4 * g + 3] = reinterpret_cast<const uint4*>(pg + p.hs + 2);
        on[4 * g] = on[4 * g + 1] = on[4 * g + 2] = on[4 * g + 3] = g < G;
      }
      ll_wait_units<4 * kSMaxChunks>(ptr, on, u, seq, p.err);
      if (G == 1) {
        v[k] = ll_f4(u[0], u[1]);
      } else {
        double M = -INFINITY;
#pragma unroll
        for (int g = 0; g < kSMaxChunks; ++g)
          if (g < G) M = fmax(M, (double)__uint_as_float(u[4 * g + 2].x));
        double wgt[kSMaxChunks], den = 0.0;
#pragma unroll
        for (int g = 0; g < kSMaxChunks; ++g) {
          wgt[g] = 0.0;
          if (g < G) {
            const double sg = __hiloint2double((int)u[4 * g + 3].x, (int)u[4 * g + 2].z);
            wgt[g] = exp((double)__uint_as_float(u[4 * g + 2].x) - M) * sg;
            den += wgt[g];
          }
        }
        double ax = 0.0, ay = 0.0, az = 0.0, aw = 0.0;
#pragma unroll
        for (int g = 0; g < kSMaxChunks; ++g) {
          if (g < G) {
            const float4 o = ll_f4(u[4 * g], u[4 * g + 1]);
            const double wg = wgt[g] / den;
            ax += (double)o.x * wg;
            ay += (double)o.y * wg;
            az += (double)o.z * wg;
            aw += (double)o.w * wg;
          }
        }
        v[k] = make_float4((float)ax, (float)ay, (float)az, (float)aw);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSMaxV4D; ++k) {
    const int j = threadIdx.x + k * kSThreads;
    if (j < n4) XV::store(xs, n4, j, v[k]);
  }
  __syncthreads();
}

// ---- one GEMV phase, every warp from its own ring --------------------------------------------
// Code size matters here: the CTA's warps run different phases of this kernel at the same
// time, and a warp that executes a cold instruction path once per phase pays an L2 round
// trip per cache line.  One GEMV body for all five streams, epilogue chosen at run time.
__device__ __noinline__ void s_gemv(const StreamParams& p, SWarp& w, const int kind, const int l, const int pos,
                                    const uint32_t seq, const unsigned char* xs, const SShared& sh, float& bv,
                                    int& bi) {
  typedef XVec<true> XV;
  const int lane = threadIdx.x & 31;
  const int pa = sh.rng[kind][threadIdx.x >> 5][0], pb = sh.rng[kind][threadIdx.x >> 5][1];
  if (pa >= pb) return;
  const int n = sh.n[kind], n4 = n >> 2;
  const int per = sh.per[kind];                       // > 0: stages hold `per` whole row pairs
  const int nseg = per > 0 ? 1 : (n + kSSeg - 1) / kSSeg;
  const int own0 = 2 * sh.cta[SP_WO][0];
  const int ncta = (int)gridDim.x;
  const float4* st = nullptr;
  int q = 0, np = 0;                                  // pair inside the stage, pairs in the stage
#pragma unroll 1
  for (int pair = pa; pair < pb; ++pair) {
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll 1
    for (int sg = 0; sg < nseg; ++sg) {
      if (per == 0 || q == 0) {
        st = s_acquire(w);
        const int left = pb - pair;
        np = per < left ? per : left;
      }
      const float4* r0 = per > 0 ? st + (size_t)q * 2 * n4 : st;
      const float4* r1 = per > 0 ? r0 + n4 : st + kSSeg / 4;
      const int j0 = sg * (kSSeg / 4) + lane;
#pragma unroll
      for (int u = 0; u < kSSeg / 128; ++u) {
        const int idx = j0 + u * 32;
        if (idx < n4) {
          const float4 a = r0[lane + u * 32], b = r1[lane + u * 32];
          double xv[4];
          XV::load(xs, n4, idx, xv);
          double c0 = acc[0][u & 1], c1 = acc[1][u & 1];
          c0 = fma((double)a.x, xv[0], c0);
          c1 = fma((double)b.x, xv[0], c1);
          c0 = fma((double)a.y, xv[1], c0);
          c1 = fma((double)b.y, xv[1], c1);
          c0 = fma((double)a.z, xv[2], c0);
          c1 = fma((double)b.z, xv[2], c1);
          c0 = fma((double)a.w, xv[3], c0);
          c1 = fma((double)b.w, xv[3], c1);
          acc[0][u & 1] = c0;
          acc[1][u & 1] = c1;
        }
      }
      if (per == 0) {
        s_release(p, sh, w);
      } else if (++q == np) {
        q = 0;
        s_release(p, sh, w);
      }
    }
    const double d0 = warp_sum_f64(acc[0][0] + acc[0][1]);
    const double d1 = warp_sum_f64(acc[1][0] + acc[1][1]);
    // epilogue: the whole warp holds the same two row sums
    const float s0 = (float)d0, s1 = (float)d1;
    const int r = 2 * pair;
    if (kind == SP_QKV) {
      if (lane == 0) {
        const int seg = r >= 2 * p.D ? 2 : (r >= p.D ? 1 : 0);
        const int i = r - seg * p.D;
        const int h = i / p.hs, c = i - h * p.hs;
        const size_t row = (size_t)l * p.kv_layer + ((size_t)h * p.steps + pos) * p.hs + c;
        if (seg == 2) {  // value row pair, llama2.ts:240
          st_gpu_u4(p.v_ll + i, __float_as_uint(s0), seq, __float_as_uint(s1), seq);
          p.vc[row] = s0;
          p.vc[row + 1] = s1;
        } else {         // RoPE, llama2.ts:223-236
          const double fr = (double)__ldg(p.fcr + (size_t)pos * (p.hs / 2) + c / 2);
          const double fi = (double)__ldg(p.fci + (size_t)pos * (p.hs / 2) + c / 2);
          const float o0 = (float)((double)s0 * fr - (double)s1 * fi);
          const float o1 = (float)((double)s0 * fi + (double)s1 * fr);
          if (seg == 0) {
            st_gpu_u4(p.q_ll + i, __float_as_uint(o0), seq, __float_as_uint(o1), seq);
          } else {
            st_gpu_u4(p.k_ll + i, __float_as_uint(o0), seq, __float_as_uint(o1), seq);
            p.kc[row] = o0;
            p.kc[row + 1] = o1;
          }
        }
      }
    } else if (kind == SP_WO || kind == SP_W2) {  // residual, llama2.ts:271, 294
      const float x0 = (float)((double)sh.xown[r - own0] + (double)s0);
      const float x1 = (float)((double)sh.xown[r + 1 - own0] + (double)s1);
      for (int c = lane; c < ncta; c += 32)
        st_gpu_u4(p.x_ll + (size_t)c * p.mbox_stride + r, __float_as_uint(x0), seq, __float_as_uint(x1), seq);
    } else if (kind == SP_W13) {                  // SwiGLU, llama2.ts:281-286
      const double hv = (double)s0;
      const float silu = (float)(hv * (1.0 / (1.0 + exp(-hv))));
      const float o = (float)((double)silu * (double)s1);
      for (int c = lane; c < ncta; c += 32)
        st_gpu_u2(p.hb_ll + (size_t)c * p.mbox_stride + pair, __float_as_uint(o), seq);
    } else if (lane == 0) {
      p.logits[r] = s0;
      p.logits[r + 1] = s1;
      argmax_consider(s0, r, bv, bi);
      argmax_consider(s1, r + 1, bv, bi);
    }
  }
}

// ---- attention: head h, time chunk g of G (llama2.ts:244-267) ----------------------------------
// The newest K/V row comes from the LL words of this step's q/k/v phase, older rows from the
// cache.  Publishes the locally normalised weighted sum, the chunk's score maximum and exp-sum.
__device__ __noinline__ void s_attention(const StreamParams& p, const int l, const int h, const int g, const int G,
                                         const int pos, const uint32_t seq_in, const uint32_t seq_out, float* sc,
                                         float* s_red, float* s_cur /* q | k | v, hs each */, SShared& sh) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hs = p.hs, hs4 = hs >> 2;
  const int n_t = pos + 1;
  const int t0 = (n_t * g) / G, t1 = (n_t * (g + 1)) / G;
  const bool has_cur = (t1 == n_t);
  // q, and the current k/v rows, from the LL words
  {
    const int units = hs >> 1;  // 16-byte units per vector
    const int which = tid / units, u = tid - which * units;
    if (which < (has_cur ? 3 : 1)) {
      const uint2* src = (which == 0 ? p.q_ll : which == 1 ? p.k_ll : p.v_ll) + (size_t)h * hs + 2 * u;
      const uint4 a = ll_wait2(src, seq_in, p.err);
      s_cur[which * hs + 2 * u] = __uint_as_float(a.x);
      s_cur[which * hs + 2 * u + 1] = __uint_as_float(a.z);
    }
  }
  __syncthreads();
  const int Gl = hs4 <= 16 ? 16 : 32;  // lanes per row
  const int subs = 32 / Gl;
  const int sub = lane / Gl, c4 = lane % Gl;
  const float* kbase = p.kc + (size_t)l * p.kv_layer + (size_t)h * p.steps * hs;
  const float* vbase = p.vc + (size_t)l * p.kv_layer + (size_t)h * p.steps * hs;
  float4 qv[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = c4 + i * Gl;
    qv[i] = c < hs4 ? reinterpret_cast<const float4*>(s_cur)[c] : f4_zero();
  }
  const double sqrt_hs = sqrt((double)hs);
  const int rows_per_pass = kSWarps * subs;
  float lmax = -INFINITY;
  for (int base = t0 + warp * subs; base < t1; base += rows_per_pass) {
    const int t = base + sub;
    const bool valid = t < t1;
    double d = 0.0;
    if (valid) {
      const bool cur = (t == pos);
      const float4* k4 = reinterpret_cast<const float4*>(kbase + (size_t)t * hs);
      const float4* kcur = reinterpret_cast<const float4*>(s_cur + hs);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int c = c4 + k * Gl;
        if (c < hs4) {
          const float4 kv = cur ? kcur[c] : ld_cg4(k4 + c);
          d = fma((double)qv[k].x, (double)kv.x, d);
          d = fma((double)qv[k].y, (double)kv.y, d);
          d = fma((double)qv[k].z, (double)kv.z, d);
          d = fma((double)qv[k].w, (double)kv.w, d);
        }
      }
    }
    for (int o = Gl >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (valid && c4 == 0) {
      const float sv = (float)(d / sqrt_hs);
      sc[t - t0] = sv;
      lmax = fmaxf(lmax, sv);
    }
  }
  lmax = warp_max_f32(lmax);
  if (lane == 0) sh.wmax[warp] = lmax;
  __syncthreads();
  float gmax = sh.wmax[0];
#pragma unroll
  for (int ww = 1; ww < kSWarps; ++ww) gmax = fmaxf(gmax, sh.wmax[ww]);
  const int nrows = t1 - t0;
  double lsum = 0.0;
  for (int t = tid; t < nrows; t += kSThreads) {
    const float ex = (float)exp((double)sc[t] - (double)gmax);
    sc[t] = ex;
    lsum += (double)ex;
  }
  lsum = warp_sum_f64(lsum);
  if (lane == 0) sh.wsum[warp] = lsum;
  __syncthreads();
  double gsum = 0.0;
#pragma unroll
  for (int ww = 0; ww < kSWarps; ++ww) gsum += sh.wsum[ww];
  for (int t = tid; t < nrows; t += kSThreads) sc[t] = (float)((double)sc[t] / gsum);
  __syncthreads();

  float4 acc[2] = {f4_zero(), f4_zero()};
  for (int t = t0 + warp * subs + sub; t < t1; t += rows_per_pass) {
    const float a = sc[t - t0];
    const bool cur = (t == pos);
    const float4* v4 = reinterpret_cast<const float4*>(vbase + (size_t)t * hs);
    const float4* vcur = reinterpret_cast<const float4*>(s_cur + 2 * hs);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = c4 + k * Gl;
      if (c < hs4) {
        const float4 vv = cur ? vcur[c] : ld_cg4(v4 + c);
        acc[k].x = fmaf(a, vv.x, acc[k].x);
        acc[k].y = fmaf(a, vv.y, acc[k].y);
        acc[k].z = fmaf(a, vv.z, acc[k].z);
        acc[k].w = fmaf(a, vv.w, acc[k].w);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    for (int o = 16; o >= Gl; o >>= 1) {
      acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, o);
      acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, o);
      acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, o);
      acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, o);
    }
    const int c = c4 + k * Gl;
    if (sub == 0 && c < hs4) reinterpret_cast<float4*>(s_red + (size_t)warp * kAttnMaxHs)[c] = acc[k];
  }
  __syncthreads();
  // the partial {out[hs], max, sum.lo, sum.hi, pad} as raw words in s_cur, then into every mailbox
  uint32_t* s_out = reinterpret_cast<uint32_t*>(s_cur);
  if (tid < hs) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < kSWarps; ++ww) s += s_red[(size_t)ww * kAttnMaxHs + tid];
    s_out[tid] = __float_as_uint(s);
  } else if (tid == hs) {
    s_out[hs] = __float_as_uint(gmax);
    s_out[hs + 1] = (uint32_t)__double2loint(gsum);
    s_out[hs + 2] = (uint32_t)__double2hiint(gsum);
    s_out[hs + 3] = 0u;
  }
  __syncthreads();
  {
    const int units = (hs + 4) >> 1;
    uint2* out = p.part_ll + ((size_t)h * kSMaxChunks + g) * p.part_stride;
    for (int c = warp; c < (int)gridDim.x; c += kSWarps)
      for (int u = lane; u < units; u += 32)
        st_gpu_u4(out + (size_t)c * p.mbox_stride + 2 * u, s_out[2 * u], seq_out, s_out[2 * u + 1], seq_out);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSThreads, 1) stream_decode_kernel(const __grid_constant__ StreamParams p) {
  extern __shared__ __align__(128) unsigned char s_smem[];
  __shared__ __align__(8) uint64_t s_bars[kSWarps][kSMaxStages];
  __shared__ SShared sh;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D, F = p.F, L = p.L;
  // dynamic smem: rings | activation vector (attention phase: scores | partial sums | q,k,v rows)
  unsigned char* xs = s_smem + (size_t)kSWarps * p.stages * kSStageFloats * 4;
  float* sc = reinterpret_cast<float*>(xs);
  float* s_red = sc + ((p.steps + 3) & ~3);
  float* s_cur = s_red + (size_t)kSWarps * kAttnMaxHs;

  SWarp w;
  w.stages = p.stages;
  w.ring = reinterpret_cast<float*>(s_smem) + (size_t)warp * p.stages * kSStageFloats;
  w.bars = &s_bars[warp][0];
  w.pol = make_l2_policy(p.evict_first != 0);
  w.c_slot = 0;
  w.c_parity = 0u;
  w.f_slot = 0;
  w.f_ph = 0;
  w.f_step = 0;
  w.f_done = false;
  if (lane == 0) {
    for (int s = 0; s < p.stages; ++s) mbar_init(w.bars + s, 1);
    mbar_fence_init();
  }
  s_tabulate(p, sh);
  __syncthreads();
  s_cursor_enter(p, sh, w);
  for (int s = 0; s < p.stages; ++s) s_fetch(p, sh, w);

  int token = ld_volatile_i32(p.ctl + CTL_HDR);
  int pos = ld_volatile_i32(p.ctl + CTL_HDR + 1);
  const int step0 = ld_volatile_i32(p.ctl + CTL_STEP);
  const int use_forced = ld_volatile_i32(p.ctl + CTL_USE_FORCED);
  const int advance = ld_volatile_i32(p.ctl + CTL_ADVANCE);
  const bool has_d_rows = sh.cta[SP_WO][0] < sh.cta[SP_WO][1];
  const uint2* my_x = p.x_ll + (size_t)blockIdx.x * p.mbox_stride;
  const uint2* my_hb = p.hb_ll + (size_t)blockIdx.x * p.mbox_stride;
  const uint2* my_am = p.am_ll + (size_t)blockIdx.x * p.mbox_stride;

#pragma unroll 1
  for (int step = 0; step < p.n_steps; ++step) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    int ev = 0;
    const bool dbg_on = p.dbg != nullptr && step + 1 == p.n_steps && threadIdx.x == 0 &&
                        (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
    auto mark = [&]() {
      if (dbg_on && ev < 1024) p.dbg[((size_t)ev * 2 + (blockIdx.x == 0 ? 0 : 1)) * 6] = gtimer_ns();
      ++ev;
    };
    mark();
    const uint32_t sbase = p.seq0 + (uint32_t)step * (uint32_t)(L + 1) * SK_KINDS;
    const int n_t = pos + 1;
    int G = n_t / 32;
    G = G < 1 ? 1 : (G > p.gmax ? p.gmax : G);
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const uint32_t sl = sbase + (uint32_t)l * SK_KINDS;
      s_gather_rms(p, my_x, sl - SK_KINDS + SK_XW2, l == 0 ? p.tok_emb + (size_t)token * D : nullptr,
                   p.rms_att + (size_t)l * D, xs, sh, nullptr);
      mark();
      s_gemv(p, w, SP_QKV, l, pos, sl + SK_QKV, xs, sh, bv, bi);
      mark();
      __syncthreads();  // xs becomes attention scratch
      for (int u = blockIdx.x; u < p.H * G; u += gridDim.x)
        s_attention(p, l, u / G, u % G, G, pos, sl + SK_QKV, sl + SK_ATT, sc, s_red, s_cur, sh);
      mark();
      if (has_d_rows) {
        s_gather_attn(p, sl + SK_ATT, G, xs);
        mark();
        s_gemv(p, w, SP_WO, l, pos, sl + SK_XWO, xs, sh, bv, bi);
        mark();
      } else {
        ev += 2;
      }
      s_gather_rms(p, my_x, sl + SK_XWO, nullptr, p.rms_ffn + (size_t)l * D, xs, sh, nullptr);
      mark();
      s_gemv(p, w, SP_W13, l, pos, sl + SK_HB, xs, sh, bv, bi);
      mark();
      if (has_d_rows) {
        s_gather_copy(p, my_hb, sl + SK_HB, F, xs);
        mark();
        s_gemv(p, w, SP_W2, l, pos, sl + SK_XW2, xs, sh, bv, bi);
        mark();
      } else {
        ev += 2;
      }
    }
    // final rmsnorm -> classifier -> argmax (llama2.ts:297-300, 364-366)
    const uint32_t sl = sbase + (uint32_t)L * SK_KINDS;
    s_gather_rms(p, my_x, sl - SK_KINDS + SK_XW2, nullptr, p.rms_final, xs, sh,
                 (blockIdx.x == 0 && step + 1 == p.n_steps) ? p.x_out : nullptr);
    s_gemv(p, w, SP_CLS, 0, pos, sl, xs, sh, bv, bi);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      argmax_consider(ov, oi, bv, bi);
    }
    if (lane == 0) {
      sh.bv[warp] = bv;
      sh.bi[warp] = bi;
    }
    __threadfence();  // this step's KV rows and logits: visible before the argmax words are
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int ww = 1; ww < kSWarps; ++ww) argmax_consider(sh.bv[ww], sh.bi[ww], bv, bi);
      if (blockIdx.x == 0) {  // owner of logits[0]: a NaN there makes the reference return 0
        const float l0 = p.logits[0];
        if (l0 != l0) {
          bv = INFINITY;
          bi = 0;
        }
      }
      sh.bv[0] = bv;
      sh.bi[0] = bi;
    }
    __syncthreads();
    if (warp == 0) {
      const float cbv = sh.bv[0];
      const int cbi = sh.bi[0];
      for (int c = lane; c < (int)gridDim.x; c += 32)
        st_gpu_u4(p.am_ll + (size_t)c * p.mbox_stride + 2 * blockIdx.x, __float_as_uint(cbv), sl, (uint32_t)cbi, sl);
      float v = -INFINITY;
      int i = 0x7fffffff;
      for (int g = lane; g < (int)gridDim.x; g += 32) {
        const uint4 a = ll_wait2(my_am + 2 * g, sl, p.err);
        argmax_consider(__uint_as_float(a.x), (int)a.z, v, i);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        argmax_consider(ov, oi, v, i);
      }
      if (lane == 0) {
        if (i == 0x7fffffff) i = 0;
        const int st = step0 + step;
        int chosen = i;
        if (use_forced) {
          const int f = p.forced[st];
          if (f >= 0) chosen = f;
        }
        if (blockIdx.x == 0) {
          p.next[0] = i;
          p.out_tokens[st] = chosen;
          if (advance) {
            p.ctl[CTL_HDR] = chosen;
            p.ctl[CTL_HDR + 1] = pos + 1;
            p.ctl[CTL_STEP] = st + 1;
          }
        }
        sh.tok = advance ? chosen : token;
        sh.pos = advance ? pos + 1 : pos;
      }
    }
    __threadfence();
    __syncthreads();
    token = sh.tok;
    pos = sh.pos;
  }
}

}  // namespace l2b
