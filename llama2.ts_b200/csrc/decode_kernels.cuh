// decode_kernels.cuh -- decode step of llama2.ts transformer() (llama2.ts:205-303)
// as five fused sm_100a kernels per layer + a classifier.
//
//   l2b_rowpair_matvec_kernel<PRO_RMS , EPI_QKV   >  rmsnorm -> wq/wk/wv matvec -> RoPE -> KV write
//   l2b_attn_decode_kernel                       scores / softmax / weighted sum, one cluster per head
//   l2b_rowpair_matvec_kernel<PRO_COPY, EPI_RESID >  wo matvec + residual
//   l2b_rowpair_matvec_kernel<PRO_RMS , EPI_SWIGLU>  rmsnorm -> w1/w3 matvec -> SiLU * gate
//   l2b_rowpair_matvec_kernel<PRO_COPY, EPI_RESID >  w2 matvec + residual
//   l2b_rowpair_matvec_kernel<PRO_RMS , EPI_LOGITS>  final rmsnorm -> wcls matvec -> device argmax
//
// All GEMV phases share one shape: a persistent grid (a multiple of the SM
// count), each CTA owning a contiguous, balanced range of ROW PAIRS of a
// row-major matrix; a warp streams both rows of a pair with 128-bit loads
// (double-buffered in registers so 2*U loads per lane are always in flight),
// multiplies against the activation vector(s) held in shared memory, reduces
// with warp shuffles and runs a pair epilogue.  Pairs are what the reference's
// element-wise stages couple: RoPE rotates (i, i+1) (llama2.ts:224-235) and
// SwiGLU couples row i of w1 with row i of w3 (llama2.ts:284-289; the upload
// interleaves those rows).
//
// NB (1,2,4,8) independent sequences share one pass over the weights (the
// small-batch path; large batches use the tcgen05 GEMM in batch_gemm.cuh).
//
// F64 = true accumulates like the reference does (JS numbers are doubles,
// llama2.ts:199-201): the product of two floats is exact in double, so a DFMA
// chain reproduces `sum += w*x` up to summation order, and the single rounding
// to float at the store makes the result bit-identical to the reference's in all
// but ~1e-8 of the elements.  B200 has a full-rate FP64 pipe, the kernels stay
// HBM-bound.  F64 = false is the plain fp32-FMA variant (kept for comparison).
//
// The first weight tile of every warp is requested BEFORE griddep_wait(): weights
// never depend on the previous kernel, so with programmatic dependent launch the
// HBM pipe stays busy across kernel boundaries.
#pragma once
#include <math.h>

#include "common.cuh"

namespace l2b {

enum { PRO_COPY = 0, PRO_RMS = 1 };
enum { EPI_QKV = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_LOGITS = 3 };

constexpr int kU = 4;          // float4 loads per lane per row per tile
constexpr int kMaxNB = 8;      // sequences sharing one weight pass
constexpr int kChunkTiles = 4; // tiles (of 32 lanes x kU float4 = 512 floats) per K-chunk of a row: the unit of the
                               // canonical summation order and of the intra-CTA split of long rows

// Tensor-parallel exchange (row-sharded projections).  Every rank
// holds replicas of the gathered vectors (x, xb, hb, logits); a kernel's epilogue stores its
// slice straight into EVERY peer's replica over NVLink (peer-mapped pointers).  The
// per-layer vectors use a flag-in-data ("LL") layout: each element is an 8-byte
// {value, sequence number} word written with ONE store, and the consumer kernel spins on
// the sequence number of exactly the elements it loads -- no fence, no ticket, no separate
// flag write on the critical path (a first version with __threadfence_system + last-CTA
// flags cost ~16 us per exchange on 2 B200s; there are 4 exchanges per layer).  Only the
// classifier -> finalize hand-off (once per step, plain logits for the host) keeps the
// fence + flag protocol.  No NCCL launch anywhere on the path.
constexpr int kMaxTp = 8;
struct TpParams {
  int rank, size;
  const int* epoch;            // device word: sequence base of the current step
  int ll_in;                   // 1: the input vector is an LL replica tagged with sequence wait_idx
  int wait_idx;                // e_in
  int out_idx;                 // e_out
  int out_off;                 // rank * slice: where this rank's results sit in the gathered vector
  int* ticket;                 // local ticket counter (last CTA publishes)
  int* err;                    // local error word (set on a wait time-out)
  int* peer_flags[kMaxTp];     // &peer.flags[e_out][rank]
  float* peer_out[kMaxTp];     // peer replicas of the output vector
  float* peer_am_val[kMaxTp];  // classifier: per-rank argmax candidates on every peer
  int* peer_am_idx[kMaxTp];
};

// Bounded spins of the exchange.  The error word `err` is followed in memory by the time-out in
// units of 2^20 clocks (option "tp_timeout_ms", default 20 s): when a peer's data does not show
// up in time the word is set and the host reports L2B_ECOMM instead of hanging the GPU.  Once it
// is set, every later wait of the step gives up after a few polls (a void step costs
// microseconds, not 129 time-outs), and l2b_tp_finalize_kernel still moves the exchange epoch on,
// so all ranks stay in lockstep and the host can re-issue the step.
__device__ __forceinline__ bool tp_spin_expired(long long t0, int* err, int& polls) {
  if ((++polls & 63) != 0) return false;
  if (ld_act_i32(err) != 0) return true;
  if (clock64() - t0 > ((long long)ld_act_i32(err + 1) << 20)) {
    atomicExch(err, 1);
    return true;
  }
  return false;
}
__device__ __forceinline__ void tp_wait_flag(const int* f, int seq, int* err) {
  const long long t0 = clock64();
  int polls = 0;
  while (ld_acquire_sys_i32(f) < seq) {
    __nanosleep(40);
    if (tp_spin_expired(t0, err, polls)) break;
  }
}

__device__ __forceinline__ uint4 ld_volatile_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_sys_u4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void st_sys_u2(void* p, uint32_t a, uint32_t b) {
  asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
// elements 4j..4j+3 of an LL replica; check: spin until all four carry sequence `seq`
__device__ __forceinline__ float4 ll_load4(const void* base, int j, bool check, int seq, int* err) {
  const uint4* b4 = reinterpret_cast<const uint4*>(base) + 2 * (size_t)j;
  uint4 a = ld_volatile_u4(b4), b = ld_volatile_u4(b4 + 1);
  if (check) {
    const long long t0 = clock64();
    int polls = 0;
    while (!((int)a.y == seq && (int)a.w == seq && (int)b.y == seq && (int)b.w == seq)) {
      if (tp_spin_expired(t0, err, polls)) break;
      a = ld_volatile_u4(b4);
      b = ld_volatile_u4(b4 + 1);
    }
  }
  return make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
}

// same with the two 16-byte words already requested (several elements' loads in flight at once:
// the spin of one element must not serialise the L2 round trips of the next ones)
__device__ __forceinline__ float4 ll_finish4(const void* base, int j, uint4 a, uint4 b, int seq, int* err) {
  const uint4* b4 = reinterpret_cast<const uint4*>(base) + 2 * (size_t)j;
  const long long t0 = clock64();
  int polls = 0;
  while (!((int)a.y == seq && (int)a.w == seq && (int)b.y == seq && (int)b.w == seq)) {
    if (tp_spin_expired(t0, err, polls)) break;
    a = ld_volatile_u4(b4);
    b = ld_volatile_u4(b4 + 1);
  }
  return make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z));
}

struct GemvParams {
  const float* W;        // [rows][n] row-major, rows even
  int rows;
  int n;                 // row length == length of the input vector
  const float* vin;      // input vectors (device activations), [B][vin_stride]
  int vin_stride;
  const float* rms_w;    // PRO_RMS: rmsnorm weight (llama2.ts:172-179)
  const float* tok_emb;  // != nullptr: layer 0, vin := tok_emb[token] (llama2.ts:211)
  const int* tokp;       // [B] device tokens
  const int* posp;       // [B] device positions
  float* x;              // [B][xdim] residual stream (EPI_RESID read-modify-write; embed target)
  int xdim;
  // EPI_QKV
  float* q;              // [B][Dq]
  float* kc;             // this layer's key cache   [B][H][steps][hs]
  float* vc;             // this layer's value cache
  const float* fcr;      // freq_cis_real [seq_len][hs/2]
  const float* fci;      // freq_cis_imag
  int Dq, hs, steps;     // Dq = rows per q/k/v segment (= dim, or dim/tp)
  long long kv_seq_stride;  // floats between two sequences inside one layer's cache
  // EPI_SWIGLU
  float* hb;             // [B][hb_stride]
  int hb_stride;
  // EPI_LOGITS
  float* logits;         // [B][V]
  int V;
  float* blk_val;        // [grid][kMaxNB] per-CTA best value
  int* blk_idx;          // [grid][kMaxNB] per-CTA best index
  int* ctl;              // host-written header: CTL_* then tok[B], pos[B]
  int* ticket;           // device-only ticket counter
  int* next;             // [B] device argmax of the last step
  const int* forced;     // [cap][B] forced next tokens (prompt), -1 = use argmax
  int* out_tokens;       // [cap][B]
  // batch slice handled by this launch
  int b0, nact, B;
  int evict_first;
  int l2_prefetch;       // bytes of this CTA's weight range pulled into L2 before griddep_wait()
  int max_local_pairs;   // upper bound of row pairs per CTA (sizes the chunk-sum scratch in shared memory)
  int split_k;           // 1: K-chunks of a pair are separate work units (few, long rows per CTA)
  long long* dbg;        // optional per-launch timeline (globaltimer ns): [launch][SM-sampled CTA][6]
  int dbg_slot;
  // software hand-over (common.cuh): wait for `sync_target` arrivals on sync_wait instead of
  // griddepcontrol.wait (nullptr: griddepcontrol.wait), arrive on sync_done at the end
  const int* sync_wait;
  int sync_target;
  int* sync_done;
  TpParams tp;           // used by the TP = true instantiations only
};

// host-written control header (ints)
enum { CTL_STEP = 0, CTL_USE_FORCED = 1, CTL_ADVANCE = 2, CTL_RESERVED = 3, CTL_HDR = 4 };

struct PairTile {
  float4 a[kU];
  float4 b[kU];
};

__device__ __forceinline__ void load_pair_tile(PairTile& t, const float4* __restrict__ w0,
                                               const float4* __restrict__ w1, int j0, int n4,
                                               uint64_t pol) {
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const int idx = j0 + u * 32;
    if (idx < n4) {
      t.a[u] = ldg_stream(w0 + idx, pol);
      t.b[u] = ldg_stream(w1 + idx, pol);
    } else {
      t.a[u] = f4_zero();
      t.b[u] = f4_zero();
    }
  }
}

// Activation vectors in shared memory.  F64: each float4 of x is kept as two
// double2 halves in two separate arrays so that consecutive lanes read
// consecutive 16-byte words (no bank conflicts).  F32: plain float4.
template <bool F64>
struct XVec;
template <>
struct XVec<true> {
  typedef double acc_t;
  static __host__ __device__ size_t bytes(int n) { return (size_t)n * 8; }
  static __device__ __forceinline__ void store(unsigned char* base, int n4, int idx, float4 v) {
    double2* A = reinterpret_cast<double2*>(base);
    double2* B = A + n4;
    A[idx] = make_double2((double)v.x, (double)v.y);
    B[idx] = make_double2((double)v.z, (double)v.w);
  }
  static __device__ __forceinline__ void load(const unsigned char* base, int n4, int idx,
                                              double (&o)[4]) {
    const double2* A = reinterpret_cast<const double2*>(base);
    const double2* B = A + n4;
    const double2 lo = A[idx], hi = B[idx];
    o[0] = lo.x; o[1] = lo.y; o[2] = hi.x; o[3] = hi.y;
  }
};
template <>
struct XVec<false> {
  typedef float acc_t;
  static __host__ __device__ size_t bytes(int n) { return (size_t)n * 4; }
  static __device__ __forceinline__ void store(unsigned char* base, int n4, int idx, float4 v) {
    reinterpret_cast<float4*>(base)[idx] = v;
  }
  static __device__ __forceinline__ void load(const unsigned char* base, int n4, int idx,
                                              float (&o)[4]) {
    const float4 v = reinterpret_cast<const float4*>(base)[idx];
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};

__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }

// argmax candidate order of llama2.ts:364-366: larger value wins, ties keep the
// lower index, NaN never wins.
__device__ __forceinline__ void argmax_consider(float v, int i, float& bv, int& bi) {
  if (v > bv || (v == bv && i < bi)) {
    bv = v;
    bi = i;
  }
}

template <int PRO, int EPI, int NB, int THREADS, bool F64, bool TP = false>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 && NB <= 2) ? 2 : 1)
l2b_rowpair_matvec_kernel(const __grid_constant__ GemvParams p) {
  constexpr int WARPS = THREADS / 32;
  constexpr int KACC = (NB == 1) ? 2 : 1;  // independent FMA chains per (row, sequence)
  typedef XVec<F64> XV;
  typedef typename XV::acc_t acc_t;

  extern __shared__ __align__(16) unsigned char smem_raw[];  // [NB] activation vectors | chunk sums | counters
  __shared__ double red_scratch[NB][WARPS];
  __shared__ float s_bv[WARPS][NB];
  __shared__ int s_bi[WARPS][NB];
  __shared__ int s_is_last;

  griddep_launch_dependents();
  const bool dbg_on = p.dbg != nullptr && threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
  long long* dbg_row = dbg_on ? p.dbg + ((size_t)p.dbg_slot * 2 + (blockIdx.x == 0 ? 0 : 1)) * 6 : nullptr;
  if (dbg_on) dbg_row[0] = gtimer_ns();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = p.n, n4 = n >> 2;
  const size_t vec_bytes = XV::bytes(n);
  const int npairs = p.rows >> 1;
  const int pair0 = (int)(((long long)npairs * blockIdx.x) / gridDim.x);
  const int pair1 = (int)(((long long)npairs * (blockIdx.x + 1)) / gridDim.x);
  const int tpp = (n4 + 32 * kU - 1) / (32 * kU);  // tiles per pair
  // Canonical summation order of a row (the same bits whatever grid, thread count, tensor-parallel
  // degree or work split computes it): lane j of a warp owns the float4 columns j, j+32, ... of the row;
  // per K-chunk of kChunkTiles tiles it runs two alternating FMA chains and adds their sum to its
  // running total IN CHUNK ORDER; the row sum is the warp's shuffle tree over the 32 lane totals.
  // Work unit: a whole row pair -- or, when a CTA owns fewer pairs than it has warps to keep busy
  // (p.split_k: w2 with its 44 KB rows leaves 2 pairs per CTA on a tensor-parallel rank of 8), one
  // K-chunk of a pair: the units are dealt to the warps round robin, a chunk's 32 lane sums go to
  // shared memory, and the warp that completes a pair adds them per lane in chunk order.
  const int nch = (tpp + kChunkTiles - 1) / kChunkTiles;
  const bool split = p.split_k != 0 && nch > 1;
  const int upp = split ? nch : 1;                 // units per pair
  const int n_units = (pair1 - pair0) * upp;
  double2* part = reinterpret_cast<double2*>(smem_raw + (size_t)NB * vec_bytes);  // [local pair][chunk][NB][32 lanes]
  int* done = reinterpret_cast<int*>(part + (size_t)p.max_local_pairs * nch * NB * 32);  // [local pair]
  const float4* W4 = reinterpret_cast<const float4*>(p.W);
  const uint64_t pol = make_l2_policy(p.evict_first != 0);

  PairTile cur, nxt;
  int unit = warp;
  int pair = 0, jt = 0, jend = 0;
  if (unit < n_units) {
    const int pl = split ? unit / upp : unit, ch = unit - pl * upp;
    pair = pair0 + pl;
    jt = split ? ch * kChunkTiles : 0;
    jend = split ? (jt + kChunkTiles < tpp ? jt + kChunkTiles : tpp) : tpp;
    const float4* w0 = W4 + (size_t)(2 * pair) * n4;
    load_pair_tile(cur, w0, w0 + n4, jt * 32 * kU + lane, n4, pol);
  }
  if (split)
    for (int i = threadIdx.x; i < pair1 - pair0; i += THREADS) done[i] = 0;
  // The HBM pipe idles between two kernels (tail of the previous one, launch, our prologue).
  // Fill that time: pull the head of this CTA's weight range into L2 now, so that the main
  // loop's next tiles are L2 hits while the stream behind them ramps up.
  if (p.l2_prefetch > 0 && lane == 0) {
    const size_t cta_bytes = (size_t)(pair1 - pair0) * 2 * n * sizeof(float);
    size_t want = (size_t)p.l2_prefetch < cta_bytes ? (size_t)p.l2_prefetch : cta_bytes;
    const size_t per = ((want / WARPS) + 15) & ~(size_t)15;
    const size_t off = (size_t)warp * per;
    if (per > 0 && off < want) {
      const size_t len = (off + per <= want) ? per : ((want - off) & ~(size_t)15);
      if (len > 0)
        prefetch_l2_bulk(reinterpret_cast<const unsigned char*>(p.W) + (size_t)pair0 * 2 * n * sizeof(float) + off,
                         (uint32_t)len);
    }
  }

  // ---- everything below may depend on the previous kernel ----
  if (dbg_on) dbg_row[1] = gtimer_ns();
  if (!TP && p.sync_wait != nullptr) {
    soft_wait(p.sync_wait, p.sync_target);
  } else {
    griddep_wait();
  }
  if (dbg_on) dbg_row[2] = gtimer_ns();
  int tp_seq = 0;
  if (TP) tp_seq = ld_act_i32(p.tp.epoch) + 1;
  const bool ll_in = TP && p.tp.ll_in != 0;   // input is an LL replica: spin on its sequence tags
  const int seq_in = tp_seq + p.tp.wait_idx;

  // prologue: build the activation vector(s) in shared memory
  {
    double ss[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) {
      ss[s] = 0.0;
      unsigned char* xs = smem_raw + (size_t)s * vec_bytes;
      if (s < p.nact) {
        const int b = p.b0 + s;
        const float* src = p.vin + (size_t)b * p.vin_stride;
        if (p.tok_emb != nullptr) src = p.tok_emb + (size_t)ld_act_i32(p.tokp + b) * n;
        const float4* src4 = reinterpret_cast<const float4*>(src);
        const bool write_x = (p.tok_emb != nullptr) && blockIdx.x == 0;
        auto consume = [&](int j, float4 v) {
          if (PRO == PRO_COPY) {
            XV::store(xs, n4, j, v);
          } else {
            // llama2.ts:174, every temporary a double like a JS number
            ss[s] += (double)v.x * (double)v.x + (double)v.y * (double)v.y +
                     (double)v.z * (double)v.z + (double)v.w * (double)v.w;
          }
          if (write_x) {  // x.set(embedding row), llama2.ts:211
            if (TP) {
              uint4* xl = reinterpret_cast<uint4*>(p.x) + 2 * (size_t)j;
              xl[0] = make_uint4(__float_as_uint(v.x), 0u, __float_as_uint(v.y), 0u);
              xl[1] = make_uint4(__float_as_uint(v.z), 0u, __float_as_uint(v.w), 0u);
            } else {
              reinterpret_cast<float4*>(p.x + (size_t)b * p.xdim)[j] = v;
            }
          }
        };
        if (ll_in) {   // tagged replica: four elements' words in flight per thread before the first spin
          constexpr int BL = 4;
          for (int j0 = threadIdx.x; j0 < n4; j0 += BL * THREADS) {
            uint4 wa[BL], wb[BL];
#pragma unroll
            for (int k = 0; k < BL; ++k) {
              const int j = j0 + k * THREADS;
              if (j < n4) {
                const uint4* b4 = reinterpret_cast<const uint4*>(p.vin) + 2 * (size_t)j;
                wa[k] = ld_volatile_u4(b4);
                wb[k] = ld_volatile_u4(b4 + 1);
              }
            }
#pragma unroll
            for (int k = 0; k < BL; ++k) {
              const int j = j0 + k * THREADS;
              if (j < n4) consume(j, ll_finish4(p.vin, j, wa[k], wb[k], seq_in, p.tp.err));
            }
          }
        } else {
          for (int j = threadIdx.x; j < n4; j += THREADS) consume(j, ld_act4(src4 + j));
        }
      } else {
        for (int j = threadIdx.x; j < n4; j += THREADS) XV::store(xs, n4, j, f4_zero());
      }
    }
    if (PRO == PRO_RMS) {
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        const double w = warp_sum_f64(ss[s]);
        if (lane == 0) red_scratch[s][warp] = w;
      }
      __syncthreads();
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        if (s < p.nact) {
          double tot = 0.0;
#pragma unroll
          for (int w = 0; w < WARPS; ++w) tot += red_scratch[s][w];  // fixed order in every CTA
          tot /= (double)n;                                         // llama2.ts:175
          tot = 1.0 / sqrt(1e-5 + tot);                             // llama2.ts:176
          unsigned char* xs = smem_raw + (size_t)s * vec_bytes;
          const int b = p.b0 + s;
          const float* src = p.vin + (size_t)b * p.vin_stride;
          if (p.tok_emb != nullptr) src = p.tok_emb + (size_t)ld_act_i32(p.tokp + b) * n;
          const float4* src4 = reinterpret_cast<const float4*>(src);
          const float4* rw4 = reinterpret_cast<const float4*>(p.rms_w);
          for (int j = threadIdx.x; j < n4; j += THREADS) {
            const float4 v = ll_in ? ll_load4(p.vin, j, false, 0, nullptr) : ld_act4(src4 + j);  // L2 hit
            const float4 w = __ldg(rw4 + j);
            float4 o;  // o[j] = weight[j] * (ss * x[j]), stored as f32 (llama2.ts:177)
            o.x = (float)((double)w.x * (tot * (double)v.x));
            o.y = (float)((double)w.y * (tot * (double)v.y));
            o.z = (float)((double)w.z * (tot * (double)v.z));
            o.w = (float)((double)w.w * (tot * (double)v.w));
            XV::store(xs, n4, j, o);
          }
        }
      }
    }
  }
  __syncthreads();

  if (dbg_on) dbg_row[3] = gtimer_ns();
  // lane s of every warp runs the epilogue of sequence b0 + s
  const bool epi_lane = lane < p.nact;
  const int eb = p.b0 + (epi_lane ? lane : 0);
  int pos = 0;
  if (EPI == EPI_QKV) pos = ld_act_i32(p.posp + eb);
  float bv = -INFINITY;
  int bi = 0x7fffffff;

  // pair epilogue, run by the lanes s < nact of ONE warp with the finished sums of "their" sequence
  auto pair_epilogue = [&](int pair, double m0, double m1) {
    const float s0 = (float)m0, s1 = (float)m1;  // xout[i] = sum, llama2.ts:201
    const int r = 2 * pair;
    if (EPI == EPI_QKV) {
      const int seg = r / p.Dq, i = r - seg * p.Dq;
      const int h = i / p.hs, c = i - h * p.hs;
      const size_t row = (size_t)eb * p.kv_seq_stride + ((size_t)h * p.steps + pos) * p.hs + c;
      if (seg == 2) {  // value row pair, llama2.ts:240
        p.vc[row] = s0;
        p.vc[row + 1] = s1;
      } else {  // RoPE, llama2.ts:224-235 (table row from the checkpoint)
        const double fr = (double)__ldg(p.fcr + (size_t)pos * (p.hs / 2) + c / 2);
        const double fi = (double)__ldg(p.fci + (size_t)pos * (p.hs / 2) + c / 2);
        const float o0 = (float)((double)s0 * fr - (double)s1 * fi);
        const float o1 = (float)((double)s0 * fi + (double)s1 * fr);
        float* dst = seg == 0 ? p.q + (size_t)eb * p.Dq + i : p.kc + row;
        dst[0] = o0;
        dst[1] = o1;
      }
    } else if (EPI == EPI_RESID) {
      // accum(x, xb2), llama2.ts:168-170,273,295
      if (TP) {
        const int gi = p.tp.out_off + r;  // index in the replicated residual stream (LL words)
        const uint4 old = ld_volatile_u4(reinterpret_cast<const uint4*>(p.x) + (gi >> 1));
        const float n0 = (float)((double)__uint_as_float(old.x) + (double)s0);
        const float n1 = (float)((double)__uint_as_float(old.z) + (double)s1);
        const uint32_t sq = (uint32_t)(tp_seq + p.tp.out_idx);
        for (int g = 0; g < p.tp.size; ++g)
          st_sys_u4(reinterpret_cast<uint4*>(p.tp.peer_out[g]) + (gi >> 1), __float_as_uint(n0), sq,
                    __float_as_uint(n1), sq);
      } else {
        float* xr = p.x + (size_t)eb * p.xdim + r;
        xr[0] = (float)((double)xr[0] + (double)s0);
        xr[1] = (float)((double)xr[1] + (double)s1);
      }
    } else if (EPI == EPI_SWIGLU) {
      // rows interleaved on upload: 2i = w1 row i, 2i+1 = w3 row i.  llama2.ts:284-289
      const double hv = (double)s0;
      const float silu = (float)(hv * (1.0 / (1.0 + exp(-hv))));
      const float hv2 = (float)((double)silu * (double)s1);
      if (TP) {
        const uint32_t sq = (uint32_t)(tp_seq + p.tp.out_idx);
        for (int g = 0; g < p.tp.size; ++g)
          st_sys_u2(reinterpret_cast<uint2*>(p.tp.peer_out[g]) + p.tp.out_off + pair, __float_as_uint(hv2), sq);
      } else {
        p.hb[(size_t)eb * p.hb_stride + pair] = hv2;
      }
    } else {
      if (TP) {
        for (int g = 0; g < p.tp.size; ++g) {
          st_relaxed_sys_f32(p.tp.peer_out[g] + p.tp.out_off + r, s0);
          st_relaxed_sys_f32(p.tp.peer_out[g] + p.tp.out_off + r + 1, s1);
        }
        argmax_consider(s0, p.tp.out_off + r, bv, bi);      // global vocabulary index
        argmax_consider(s1, p.tp.out_off + r + 1, bv, bi);
      } else {
        float* lg = p.logits + (size_t)eb * p.V + r;
        lg[0] = s0;
        lg[1] = s1;
        argmax_consider(s0, r, bv, bi);
        argmax_consider(s1, r + 1, bv, bi);
      }
    }
  };

  acc_t acc[2][NB][KACC];
  double lt[2][NB];   // this lane's running totals (chunk order)
#pragma unroll
  for (int s = 0; s < NB; ++s) {
    lt[0][s] = lt[1][s] = 0.0;
#pragma unroll
    for (int k = 0; k < KACC; ++k) acc[0][s][k] = acc[1][s][k] = (acc_t)0;
  }

  bool have = unit < n_units;
  while (have) {
    // request the next tile (same unit, or the first tile of this warp's next unit) before touching
    // the current one
    int nunit = unit, npair = pair, njt = jt + 1, njend = jend;
    if (njt == jend) {
      nunit = unit + WARPS;
      if (nunit < n_units) {
        const int pl = split ? nunit / upp : nunit, ch = nunit - pl * upp;
        npair = pair0 + pl;
        njt = split ? ch * kChunkTiles : 0;
        njend = split ? (njt + kChunkTiles < tpp ? njt + kChunkTiles : tpp) : tpp;
      }
    }
    const bool more = nunit < n_units;
    if (more) {
      const float4* w0 = W4 + (size_t)(2 * npair) * n4;
      load_pair_tile(nxt, w0, w0 + n4, njt * 32 * kU + lane, n4, pol);
    }
    {
      const int j0 = jt * 32 * kU + lane;
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int idx = j0 + u * 32;
        if (idx < n4) {
          const acc_t a0 = (acc_t)cur.a[u].x, a1 = (acc_t)cur.a[u].y, a2 = (acc_t)cur.a[u].z,
                      a3 = (acc_t)cur.a[u].w;
          const acc_t b0 = (acc_t)cur.b[u].x, b1 = (acc_t)cur.b[u].y, b2 = (acc_t)cur.b[u].z,
                      b3 = (acc_t)cur.b[u].w;
#pragma unroll
          for (int s = 0; s < NB; ++s) {
            acc_t xv[4];
            XV::load(smem_raw + (size_t)s * vec_bytes, n4, idx, xv);
            acc_t r0 = acc[0][s][u % KACC], r1 = acc[1][s][u % KACC];
            r0 = fma_t(a0, xv[0], r0);
            r1 = fma_t(b0, xv[0], r1);
            r0 = fma_t(a1, xv[1], r0);
            r1 = fma_t(b1, xv[1], r1);
            r0 = fma_t(a2, xv[2], r0);
            r1 = fma_t(b2, xv[2], r1);
            r0 = fma_t(a3, xv[3], r0);
            r1 = fma_t(b3, xv[3], r1);
            acc[0][s][u % KACC] = r0;
            acc[1][s][u % KACC] = r1;
          }
        }
      }
    }
    if ((jt + 1) % kChunkTiles == 0 || jt == tpp - 1) {  // K-chunk complete: chains -> lane totals
#pragma unroll
      for (int s = 0; s < NB; ++s) {
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int k = 0; k < KACC; ++k) {
          c0 += (double)acc[0][s][k];
          c1 += (double)acc[1][s][k];
          acc[0][s][k] = acc[1][s][k] = (acc_t)0;
        }
        lt[0][s] += c0;
        lt[1][s] += c1;
      }
    }

    if (jt == jend - 1) {  // unit complete
      const int pl = pair - pair0;
      bool run_epi = !split;
      if (split) {
        // publish this chunk's lane sums; the warp that completes the pair adds the chunks per lane in
        // chunk order (below) while the other warps keep streaming
        const int ch = unit - pl * upp;
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          part[(((size_t)pl * nch + ch) * NB + s) * 32 + lane] = make_double2(lt[0][s], lt[1][s]);
          lt[0][s] = lt[1][s] = 0.0;
        }
        __threadfence_block();
        __syncwarp();
        int old = 0;
        if (lane == 0) old = atomicAdd(done + pl, 1);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == nch - 1) {
          __threadfence_block();
          run_epi = true;
#pragma unroll
          for (int s = 0; s < NB; ++s)
            for (int c2 = 0; c2 < nch; ++c2) {
              const double2 v = part[(((size_t)pl * nch + c2) * NB + s) * 32 + lane];
              lt[0][s] += v.x;
              lt[1][s] += v.y;
            }
        }
      }
      if (run_epi) {
        double m0 = 0.0, m1 = 0.0;  // the sums of "my" sequence (lane s <-> sequence s)
#pragma unroll
        for (int s = 0; s < NB; ++s) {
          const double d0 = warp_sum_f64(lt[0][s]);
          const double d1 = warp_sum_f64(lt[1][s]);
          lt[0][s] = lt[1][s] = 0.0;
          if (lane == s) {
            m0 = d0;
            m1 = d1;
          }
        }
        if (epi_lane) pair_epilogue(pair, m0, m1);
      }
    }
    unit = nunit;
    pair = npair;
    jt = njt;
    jend = njend;
    cur = nxt;
    have = more;
  }

  if (dbg_on) dbg_row[4] = gtimer_ns();   // warp 0 of this CTA finished its rows
  if (!TP && p.sync_done != nullptr) soft_signal(p.sync_done);
  if (p.dbg != nullptr) {
    __syncthreads();
    if (dbg_on) dbg_row[5] = gtimer_ns(); // all warps of this CTA finished
  }
  if (EPI == EPI_LOGITS) {
    // device argmax (llama2.ts:364-366) + the state-machine advance of :471-504
    if (lane < NB) {
      s_bv[warp][lane] = bv;
      s_bi[warp][lane] = bi;
    }
    __syncthreads();
    if (threadIdx.x < NB) {
      const int s = threadIdx.x;
      float v = s_bv[0][s];
      int i = s_bi[0][s];
      for (int w = 1; w < WARPS; ++w) argmax_consider(s_bv[w][s], s_bi[w][s], v, i);
      p.blk_val[blockIdx.x * kMaxNB + s] = v;
      p.blk_idx[blockIdx.x * kMaxNB + s] = i;
      __threadfence();
    }
    if (TP) __threadfence_system();  // the logits slice stored into the peers
    __syncthreads();
    if (threadIdx.x == 0) {
      const int t = atomicAdd(p.ticket, 1);
      s_is_last = (t == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (TP) {
      // rank-local candidate -> every peer; l2b_tp_finalize_kernel picks the global first maximum
      if (s_is_last && warp == 0) {
        __threadfence();
        float v = -INFINITY;
        int i = 0x7fffffff;
        for (int g = lane; g < (int)gridDim.x; g += 32)
          argmax_consider(ld_act(p.blk_val + g * kMaxNB), ld_act_i32(p.blk_idx + g * kMaxNB), v, i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, v, o);
          const int oi = __shfl_xor_sync(0xffffffffu, i, o);
          argmax_consider(ov, oi, v, i);
        }
        if (ld_act_i32(p.tp.err) != 0) i = -1;  // poison: every rank learns that this step is void
        if (lane < p.tp.size) {
          st_relaxed_sys_f32(p.tp.peer_am_val[lane] + p.tp.rank, v);
          asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p.tp.peer_am_idx[lane] + p.tp.rank), "r"(i)
                       : "memory");
        }
        __threadfence_system();
        __syncwarp();
        if (lane < p.tp.size) st_release_sys_i32(p.tp.peer_flags[lane], tp_seq + p.tp.out_idx);
        if (lane == 0) *p.ticket = 0;
      }
    } else if (s_is_last && warp == 0) {
      __threadfence();
      const int B = p.B;
      const int step = ld_act_i32(p.ctl + CTL_STEP);
      const int use_forced = ld_act_i32(p.ctl + CTL_USE_FORCED);
      const int advance = ld_act_i32(p.ctl + CTL_ADVANCE);
      int* tok = p.ctl + CTL_HDR;
      int* posv = tok + B;
      for (int s = 0; s < p.nact; ++s) {
        float v = -INFINITY;
        int i = 0x7fffffff;
        for (int g = lane; g < (int)gridDim.x; g += 32)
          argmax_consider(ld_act(p.blk_val + g * kMaxNB + s), ld_act_i32(p.blk_idx + g * kMaxNB + s),
                          v, i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, v, o);
          const int oi = __shfl_xor_sync(0xffffffffu, i, o);
          argmax_consider(ov, oi, v, i);
        }
        if (lane == 0) {
          const int b = p.b0 + s;
          const float l0 = ld_act(p.logits + (size_t)b * p.V);
          if (i == 0x7fffffff || l0 != l0) i = 0;  // all-NaN / NaN at index 0: reduce() keeps 0
          int chosen = i;
          if (use_forced) {
            const int f = p.forced[(size_t)step * B + b];
            if (f >= 0) chosen = f;  // prompt forcing, llama2.ts:471-473
          }
          p.next[b] = i;
          p.out_tokens[(size_t)step * B + b] = chosen;
          if (advance) {
            tok[b] = chosen;  // token = next; pos++  (llama2.ts:496,504)
            posv[b] = posv[b] + 1;
          }
        }
      }
      if (lane == 0) {
        if (advance && p.b0 + p.nact == B) p.ctl[CTL_STEP] = step + 1;
        *p.ticket = 0;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Attention for one layer, batch of independent sequences (llama2.ts:244-267).
// grid = (CS, H, B) with clusters of CS CTAs along x: the CS CTAs of a cluster
// split the time steps 0..pos of one (sequence, head), so that a single decode
// step still spreads over the chip.  The cluster exchanges its maxima and sums
// through distributed shared memory, which keeps the reference's exact two-pass
// softmax (global max, f32-rounded exponentials, f64 sum, f32-rounded
// probabilities) instead of an online-softmax approximation.
//
// K and V tiles of the head-major cache ([H][steps][hs], so one head's rows are a
// single contiguous run) arrive through a 4-stage ring of 1-D bulk async copies
// (cp.async.bulk -> UBLKCP) signalled on mbarriers; V tiles are already in flight
// while the softmax runs.
constexpr int kAttnThreads = 256;
constexpr int kAttnWarps = kAttnThreads / 32;
constexpr int kAttnStages = 4;
constexpr int kAttnStageBytes = 16384;       // ring stage of the many-CTA (batched) launches
constexpr int kAttnStageBytesBig = 32768;    // batch-1 launches (one cluster per head owns its SMs): 96 KB of
                                             // cache rows in flight per CTA; at pos ~1800 the 4 x 16 KB ring capped the
                                             // K and V passes at ~4 TB/s over 128 CTAs
constexpr int kAttnMaxHs = 256;

struct AttnParams {
  const float* q;    // [B][q_stride]   rotated queries
  const float* kc;   // layer base, [B][H][steps][hs]
  const float* vc;
  float* xb;         // [B][xb_stride]  attention output
  const int* posp;   // [B]
  int H, hs, steps;
  long long kv_b_stride;            // floats between the caches of consecutive batch entries (0: all
                                    // entries read ONE sequence's cache -- prompt prefill)
  int q_stride, xb_stride, xb_off;  // xb_off: column offset of head 0 (tensor-parallel slice)
  int tileT;         // time steps per ring stage
  int stage_bytes;   // bytes per ring stage (kAttnStageBytes or kAttnStageBytesBig)
  int sc_cap;        // floats reserved for scores per CTA
  // tensor-parallel all-gather of the output slice (nullptr when tp_size == 1)
  float* peer_xb[kMaxTp];  // peers' xb replicas (including our own)
  int tp_size;
  const int* tp_epoch;     // device word: sequence base of the current step
  int tp_out_idx;          // exchange index of this all-gather

  // batched tensor-core path: also emit the TF32 hi/lo split of the output (input of the wo GEMM)
  float* xh;
  float* xl;
  int x_npad;
  int nbatch;        // l2b_attn_warp_kernel: number of batch entries
  // attention is latency-bound and leaves HBM idle: meanwhile pull the NEXT kernel's weights
  // (wo of this layer) into L2
  const unsigned char* pf_ptr;
  long long pf_bytes;
};

__global__ void __launch_bounds__(kAttnThreads, 1) l2b_attn_decode_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(128) unsigned char attn_smem_raw[];
  float* ring = reinterpret_cast<float*>(attn_smem_raw);                      // kAttnStages * stage
  float* sc = ring + (size_t)kAttnStages * (p.stage_bytes / 4);            // sc_cap floats
  __shared__ __align__(8) uint64_t full_bar[kAttnStages];
  __shared__ __align__(8) uint64_t empty_bar[kAttnStages];
  __shared__ float s_red[kAttnWarps][kAttnMaxHs];
  __shared__ float s_wmax[kAttnWarps];
  __shared__ double s_wsum[kAttnWarps];
  __shared__ float c_out[kAttnMaxHs];   // read by rank 0 through DSMEM
  __shared__ float c_max;               // read by every rank through DSMEM
  __shared__ double c_sum;

  griddep_launch_dependents();

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank(), CS = cluster_nctarank();
  const int h = blockIdx.y, b = blockIdx.z;
  const int hs = p.hs, hs4 = hs >> 2;

  if (tid == 0) {
    for (int s = 0; s < kAttnStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kAttnWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  griddep_wait();

  const int pos = ld_act_i32(p.posp + b);
  const int n_t = pos + 1;
  const int chunk = (n_t + (int)CS - 1) / (int)CS;
  const int t0 = (int)rank * chunk;
  int nT = n_t - t0;
  nT = nT < 0 ? 0 : (nT > chunk ? chunk : nT);
  const int tileT = p.tileT;
  const int nTiles = (nT + tileT - 1) / tileT;
  const int total = 2 * nTiles;  // K tiles then V tiles
  const size_t head_off = (size_t)b * (size_t)p.kv_b_stride + ((size_t)h * p.steps) * hs;
  const float* kbase = p.kc + head_off + (size_t)t0 * hs;
  const float* vbase = p.vc + head_off + (size_t)t0 * hs;
  const int stage_floats = p.stage_bytes / 4;

  auto issue = [&](int j) {
    const int s = j % kAttnStages;
    if (j >= kAttnStages) mbar_wait(&empty_bar[s], ((j / kAttnStages) - 1) & 1);
    const int tile = j < nTiles ? j : j - nTiles;
    const float* base = j < nTiles ? kbase : vbase;
    const int tt0 = tile * tileT;
    const int cnt = min(tileT, nT - tt0);
    const uint32_t bytes = (uint32_t)cnt * hs * 4u;
    mbar_arrive_expect_tx(&full_bar[s], bytes);
    bulk_g2s(ring + (size_t)s * stage_floats, base + (size_t)tt0 * hs, bytes, &full_bar[s]);
  };
  if (tid == 0)
    for (int j = 0; j < kAttnStages - 1 && j < total; ++j) issue(j);
  if (p.pf_bytes > 0 && tid == 0) {
    // behind our own K/V requests in the copy queue: one slice of the next kernel's weights
    const long long n_cta = (long long)gridDim.x * gridDim.y * gridDim.z;
    const long long me = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const long long per = ((p.pf_bytes / n_cta) + 15) & ~15LL;
    long long off = me * per;
    const long long end = off + per < p.pf_bytes ? off + per : p.pf_bytes;
    for (; off + 16 <= end; off += 65536) {
      const long long len = (end - off < 65536 ? end - off : 65536) & ~15LL;
      if (len > 0) prefetch_l2_bulk(p.pf_ptr + off, (uint32_t)len);
    }
  }

  // lane layout: G lanes cover one cache row (hs floats) as float4s
  const int G = hs4 <= 16 ? 16 : 32;
  const int subs = 32 / G;
  const int sub = lane / G, c4 = lane % G;
  float4 qv[2];
  {
    const float4* q4 = reinterpret_cast<const float4*>(p.q + (size_t)b * p.q_stride + (size_t)h * hs);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = c4 + i * G;
      qv[i] = c < hs4 ? ld_act4(q4 + c) : f4_zero();
    }
  }
  const double sqrt_hs = sqrt((double)hs);
  const int rows_per_pass = kAttnWarps * subs;

  // ---- pass 1: scores (llama2.ts:249-254), f64 dot like the reference ----
  float lmax = -INFINITY;
  for (int i = 0; i < nTiles; ++i) {
    if (tid == 0 && i + kAttnStages - 1 < total) issue(i + kAttnStages - 1);
    __syncwarp();
    const int s = i % kAttnStages;
    mbar_wait(&full_bar[s], (i / kAttnStages) & 1);
    const float4* st4 = reinterpret_cast<const float4*>(ring + (size_t)s * stage_floats);
    const int tt0 = i * tileT;
    const int cnt = min(tileT, nT - tt0);
    for (int base = warp * subs; base < cnt; base += rows_per_pass) {
      const int tt = base + sub;
      const bool valid = tt < cnt;
      double d = 0.0;
      if (valid) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int c = c4 + k * G;
          if (c < hs4) {
            const float4 kv = st4[(size_t)tt * hs4 + c];
            d = fma((double)qv[k].x, (double)kv.x, d);
            d = fma((double)qv[k].y, (double)kv.y, d);
            d = fma((double)qv[k].z, (double)kv.z, d);
            d = fma((double)qv[k].w, (double)kv.w, d);
          }
        }
      }
      for (int o = G >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (valid && c4 == 0) {
        const float sv = (float)(d / sqrt_hs);  // scope / Math.sqrt(head_size)
        sc[tt0 + tt] = sv;
        lmax = fmaxf(lmax, sv);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  // ---- softmax over the whole cluster (llama2.ts:181-194) ----
  lmax = warp_max_f32(lmax);
  if (lane == 0) s_wmax[warp] = lmax;
  __syncthreads();
  if (tid == 0) {
    float m = s_wmax[0];
    for (int w = 1; w < kAttnWarps; ++w) m = fmaxf(m, s_wmax[w]);
    c_max = m;
  }
  cluster_sync_all();
  float gmax = -INFINITY;
  for (uint32_t r = 0; r < CS; ++r) gmax = fmaxf(gmax, dsmem_ld_f32(dsmem_addr(&c_max, r)));
  double lsum = 0.0;
  for (int t = tid; t < nT; t += kAttnThreads) {
    const float e = (float)exp((double)sc[t] - (double)gmax);  // Math.exp, stored as f32
    sc[t] = e;
    lsum += (double)e;
  }
  lsum = warp_sum_f64(lsum);
  if (lane == 0) s_wsum[warp] = lsum;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kAttnWarps; ++w) s += s_wsum[w];
    c_sum = s;
  }
  cluster_sync_all();
  double gsum = 0.0;
  for (uint32_t r = 0; r < CS; ++r) gsum += dsmem_ld_f64(dsmem_addr(&c_sum, r));
  for (int t = tid; t < nT; t += kAttnThreads) sc[t] = (float)((double)sc[t] / gsum);
  __syncthreads();

  // ---- pass 2: weighted sum of values (llama2.ts:257-265) ----
  float4 acc[2] = {f4_zero(), f4_zero()};
  for (int i = nTiles; i < total; ++i) {
    if (tid == 0 && i + kAttnStages - 1 < total) issue(i + kAttnStages - 1);
    __syncwarp();
    const int s = i % kAttnStages;
    mbar_wait(&full_bar[s], (i / kAttnStages) & 1);
    const float4* st4 = reinterpret_cast<const float4*>(ring + (size_t)s * stage_floats);
    const int tt0 = (i - nTiles) * tileT;
    const int cnt = min(tileT, nT - tt0);
    for (int tt = warp * subs + sub; tt < cnt; tt += rows_per_pass) {
      const float a = sc[tt0 + tt];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int c = c4 + k * G;
        if (c < hs4) {
          const float4 vv = st4[(size_t)tt * hs4 + c];
          acc[k].x = fmaf(a, vv.x, acc[k].x);
          acc[k].y = fmaf(a, vv.y, acc[k].y);
          acc[k].z = fmaf(a, vv.z, acc[k].z);
          acc[k].w = fmaf(a, vv.w, acc[k].w);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
  // fold the sub-row groups of a warp, then the warps, then the cluster
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    for (int o = 16; o >= G; o >>= 1) {
      acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, o);
      acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, o);
      acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, o);
      acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, o);
    }
    const int c = c4 + k * G;
    if (sub == 0 && c < hs4) reinterpret_cast<float4*>(&s_red[warp][0])[c] = acc[k];
  }
  __syncthreads();
  if (tid < hs) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) s += s_red[w][tid];
    c_out[tid] = s;
  }
  cluster_sync_all();
  if (rank == 0 && tid < hs) {
    float s = 0.f;
    for (uint32_t r = 0; r < CS; ++r) s += dsmem_ld_f32(dsmem_addr(&c_out[tid], r));
    const size_t o = (size_t)b * p.xb_stride + p.xb_off + (size_t)h * hs + tid;
    if (p.xh != nullptr) {
      const float hi = __uint_as_float((__float_as_uint(s) + 0x1000u) & 0xFFFFE000u);
      const int j = (int)(h * hs + tid);  // column of this element; rows = batch entries
      const size_t xi = ((size_t)(j >> 5) * p.x_npad + b) * 32 + ((((j >> 2) & 7) ^ (b & 7)) << 2) + (j & 3);
      p.xh[xi] = hi;
      p.xl[xi] = s - hi;
    }
    if (p.tp_size <= 1) {
      p.xb[o] = s;
    } else {
      const uint32_t sq = (uint32_t)(ld_act_i32(p.tp_epoch) + 1 + p.tp_out_idx);
      for (int r = 0; r < p.tp_size; ++r)
        st_sys_u2(reinterpret_cast<uint2*>(p.peer_xb[r]) + o, __float_as_uint(s), sq);
    }
  }
  cluster_sync_all();  // keep every CTA's shared memory alive until rank 0 has read it
}

// Batched attention: ONE WARP per (sequence, head) (llama2.ts:244-267).  With hundreds of
// independent (sequence, head) pairs there is no need to split one head over a cluster: the
// cluster kernel above spends ~8 us of latency per CTA (mbarrier setup, TMA round trips, four
// cluster barriers) on ~70 KB of cache, which made attention 19 % of a 256-sequence step.
// Here a warp streams its head's K rows then V rows with coalesced 512-byte loads, several
// rows in flight, no block-level synchronisation; the softmax keeps the reference's two-pass
// form (global max, f32-rounded exponentials, f64 sum).  Scores live in shared memory.
constexpr int kAttnWarpThreads = 256;
constexpr int kAttnWarpRows = 8;   // cache rows in flight per warp (head_size <= 128: one float4 per lane)
__global__ void __launch_bounds__(kAttnWarpThreads) l2b_attn_warp_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(16) float attn_warp_sc[];  // [warps][sc_cap]
  griddep_launch_dependents();
  griddep_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * (kAttnWarpThreads / 32) + warp;  // (b, h)
  if (item >= p.nbatch * p.H) return;
  const int b = item / p.H, h = item - b * p.H;
  const int hs = p.hs, hs4 = hs >> 2;
  const bool on = lane < hs4;
  float* sc = attn_warp_sc + (size_t)warp * p.sc_cap;
  const int pos = ld_act_i32(p.posp + b);
  const int n_t = pos + 1;
  const size_t head_off = (size_t)b * (size_t)p.kv_b_stride + ((size_t)h * p.steps) * hs;
  const float4* k4 = reinterpret_cast<const float4*>(p.kc + head_off) + lane;
  const float4* v4 = reinterpret_cast<const float4*>(p.vc + head_off) + lane;
  const float4 qv = on ? ld_act4(reinterpret_cast<const float4*>(p.q + (size_t)b * p.q_stride + (size_t)h * hs) + lane)
                       : f4_zero();
  const double sqrt_hs = sqrt((double)hs);

  // pass 1: scores
  float lmax = -INFINITY;
  for (int t0 = 0; t0 < n_t; t0 += kAttnWarpRows) {
    float4 kr[kAttnWarpRows];
#pragma unroll
    for (int u = 0; u < kAttnWarpRows; ++u)
      kr[u] = (on && t0 + u < n_t) ? __ldcg(k4 + (size_t)(t0 + u) * hs4) : f4_zero();
#pragma unroll
    for (int u = 0; u < kAttnWarpRows; ++u) {
      double d = fma((double)qv.x, (double)kr[u].x, 0.0);
      d = fma((double)qv.y, (double)kr[u].y, d);
      d = fma((double)qv.z, (double)kr[u].z, d);
      d = fma((double)qv.w, (double)kr[u].w, d);
      d = warp_sum_f64(d);
      if (t0 + u < n_t) {
        const float sv = (float)(d / sqrt_hs);
        if (lane == 0) sc[t0 + u] = sv;
        lmax = fmaxf(lmax, sv);
      }
    }
  }
  __syncwarp();
  // softmax (llama2.ts:181-194)
  double lsum = 0.0;
  for (int t = lane; t < n_t; t += 32) {
    const float e = (float)exp((double)sc[t] - (double)lmax);
    sc[t] = e;
    lsum += (double)e;
  }
  lsum = warp_sum_f64(lsum);
  __syncwarp();
  for (int t = lane; t < n_t; t += 32) sc[t] = (float)((double)sc[t] / lsum);
  __syncwarp();
  // pass 2: weighted sum of the value rows
  float4 acc = f4_zero();
  for (int t0 = 0; t0 < n_t; t0 += kAttnWarpRows) {
    float4 vr[kAttnWarpRows];
#pragma unroll
    for (int u = 0; u < kAttnWarpRows; ++u)
      vr[u] = (on && t0 + u < n_t) ? __ldcg(v4 + (size_t)(t0 + u) * hs4) : f4_zero();
#pragma unroll
    for (int u = 0; u < kAttnWarpRows; ++u) {
      const float a = (t0 + u < n_t) ? sc[t0 + u] : 0.f;
      acc.x = fmaf(a, vr[u].x, acc.x);
      acc.y = fmaf(a, vr[u].y, acc.y);
      acc.z = fmaf(a, vr[u].z, acc.z);
      acc.w = fmaf(a, vr[u].w, acc.w);
    }
  }
  if (on) {
    const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
    const size_t o = (size_t)b * p.xb_stride + p.xb_off + (size_t)h * hs + 4 * lane;
    *reinterpret_cast<float4*>(p.xb + o) = acc;
    if (p.xh != nullptr) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = h * hs + 4 * lane + e;
        const size_t xi = ((size_t)(j >> 5) * p.x_npad + b) * 32 + ((((j >> 2) & 7) ^ (b & 7)) << 2) + (j & 3);
        const float hi = __uint_as_float((__float_as_uint(vals[e]) + 0x1000u) & 0xFFFFE000u);
        p.xh[xi] = hi;
        p.xl[xi] = vals[e] - hi;
      }
    }
  }
}

// Tensor-parallel step epilogue (one warp): waits for every rank's classifier slice, picks
// the global argmax (llama2.ts:364-366: first maximum wins -> lowest global index), runs the
// state machine of llama2.ts:471-504 on this rank's replica of the control block and moves
// the exchange epoch on.  Every rank computes the same token.
struct TpFinalParams {
  int size, n_exchanges;
  int* epoch;
  const int* wait_flags;   // local flags[last exchange][0..size)
  int wait_idx;
  const float* am_val;     // [size] candidates written by the ranks
  const int* am_idx;
  const float* logits;     // local replica of the gathered logits
  int* ctl;
  int* next;
  const int* forced;
  int* out_tokens;
  int* err;
};
__global__ void l2b_tp_finalize_kernel(const __grid_constant__ TpFinalParams p) {
  griddep_launch_dependents();
  griddep_wait();
  const int lane = threadIdx.x;
  const int base = ld_act_i32(p.epoch);
  if (lane < p.size) tp_wait_flag(p.wait_flags + lane, base + 1 + p.wait_idx, p.err);
  if (lane < p.size && ld_act_i32(p.am_idx + lane) == -1) atomicExch(p.err, 1);  // a peer timed out
  __syncwarp();
  if (lane == 0 && ld_act_i32(p.err) == 0) {   // a timed-out step changes no host-visible state
    float v = -INFINITY;
    int i = 0x7fffffff;
    for (int g = 0; g < p.size; ++g) argmax_consider(ld_act(p.am_val + g), ld_act_i32(p.am_idx + g), v, i);
    const float l0 = ld_act(p.logits);
    if (i == 0x7fffffff || l0 != l0) i = 0;
    const int step = p.ctl[CTL_STEP];
    int chosen = i;
    if (p.ctl[CTL_USE_FORCED]) {
      const int f = p.forced[step];
      if (f >= 0) chosen = f;
    }
    p.next[0] = i;
    p.out_tokens[step] = chosen;
    if (p.ctl[CTL_ADVANCE]) {
      p.ctl[CTL_HDR] = chosen;
      p.ctl[CTL_HDR + 1] = p.ctl[CTL_HDR + 1] + 1;
      p.ctl[CTL_STEP] = step + 1;
    }
  }
  if (lane == 0) {   // the epoch moves on even after a time-out: ranks stay in lockstep
    *p.epoch = base + p.n_exchanges;
  }
}

}  // namespace l2b
