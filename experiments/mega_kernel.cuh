// mega_kernel.cuh -- the whole batch-1 decode step (llama2.ts:205-303), and the greedy
// loop around it (llama2.ts:465-508), as ONE persistent cooperative kernel.
//
// Why: with one kernel per fused op, every kernel boundary costs ~6 us on B200 even with
// programmatic dependent launch (launch latency, activation load + rmsnorm prologue,
// pipeline ramp, tail) -- 161 boundaries per Llama-2-7B token, ~20 % of the step, and
// nearly all of the step for the small models.  Here the 148 CTAs stay resident, phases
// are separated by a grid barrier (~1 us), and -- the point -- every warp requests the
// first weight tile of the NEXT phase before it arrives at the barrier: weights never
// depend on activations, so the HBM pipe keeps streaming through the barrier and the
// prologue of the next phase.
//
// Phases per layer:  QKV (rmsnorm, q/k/v rows, RoPE, KV write) | attention (one CTA per
// head, exact two-pass softmax) | wo + residual | rmsnorm, w1/w3, SwiGLU | w2 + residual;
// then the classifier + argmax + state advance.  Arithmetic is the same as in
// decode_kernels.cuh (fp64 accumulation, reference rounding points).
#pragma once
#include <math.h>

#include "../llama2.ts_b200/csrc/common.cuh"
#include "../llama2.ts_b200/csrc/decode_kernels.cuh"

namespace l2b {

// 256 threads with up to 255 registers each: two 8-deep register tiles (2 x 64 registers) per
// thread keep 128 KB per SM in flight, the same as the 512-thread GEMV kernels, without the
// spills a 128-register cap forces once the whole step lives in one kernel.
constexpr int kMegaThreads = 256;
constexpr int kMegaWarps = kMegaThreads / 32;
constexpr int kMU = 8;  // float4 loads per lane per row per tile

struct MegaTile {
  float4 a[kMU];
  float4 b[kMU];
};
__device__ __forceinline__ void mega_load_tile(MegaTile& t, const float4* __restrict__ w0,
                                               const float4* __restrict__ w1, int j0, int n4, uint64_t pol) {
#pragma unroll
  for (int u = 0; u < kMU; ++u) {
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

struct MegaParams {
  int D, F, L, H, hs, V, steps;
  const float *tok_emb, *rms_att, *wqkv, *wo, *rms_ffn, *w13, *w2, *rms_final, *fcr, *fci, *wcls;
  float *x, *xb, *q, *hb, *logits, *kc, *vc;
  long long kv_layer;      // floats per layer of one cache (H * steps * hs)
  int* ctl;                // CTL_* header, tok[1], pos[1]
  int* next;
  const int* forced;
  int* out_tokens;
  float* blk_val;
  int* blk_idx;
  unsigned* bar;           // [0] arrival count, [1] generation
  int n_steps;
  int evict_first;
};

// activations are rewritten by other CTAs between phases: always read them from L2
__device__ __forceinline__ float4 ld_cg4(const float4* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ float ld_cg(const float* p) {
  float r;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ int ld_cg_i32(const int* p) {
  int r;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}

// sense-reversing grid barrier (all CTAs co-resident: cooperative launch)
__device__ __forceinline__ void grid_sync(unsigned* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned gen;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
    __threadfence();
    const unsigned arrived = atomicAdd(bar, 1u);
    if (arrived == gridDim.x - 1) {
      bar[0] = 0u;
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1u) : "memory");
    } else {
      unsigned g;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(bar + 1) : "memory");
      } while (g == gen);
    }
    __threadfence();
  }
  __syncthreads();
}

struct MegaPhase {
  const float* W;
  int rows, n;
};

// First weight tile of a phase for this warp, requested before the preceding barrier.  It is
// parked in shared memory by cp.async (LDGSTS) rather than in registers: the tile then is not
// live across the barrier and the next prologue (whose fp64 sqrt/divide are real calls and
// would spill the tile around them).  stage: [warp][kMU][2 rows][32 lanes] float4.
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void mega_prefetch(const MegaPhase& ph, float4* stage) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n4 = ph.n >> 2, npairs = ph.rows >> 1;
  const int pair0 = (int)(((long long)npairs * blockIdx.x) / gridDim.x);
  const int pair1 = (int)(((long long)npairs * (blockIdx.x + 1)) / gridDim.x);
  const int pair = pair0 + warp;
  if (pair < pair1) {
    const float4* w0 = reinterpret_cast<const float4*>(ph.W) + (size_t)(2 * pair) * n4;
    float4* st = stage + (size_t)warp * (kMU * 2 * 32) + lane;
#pragma unroll
    for (int u = 0; u < kMU; ++u) {
      const int idx = lane + u * 32;
      if (idx < n4) {
        cp_async16(st + (u * 2) * 32, w0 + idx);
        cp_async16(st + (u * 2 + 1) * 32, w0 + n4 + idx);
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void mega_take_prefetched(MegaTile& cur, const float4* stage, int n4) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const float4* st = stage + (size_t)warp * (kMU * 2 * 32) + lane;
#pragma unroll
  for (int u = 0; u < kMU; ++u) {
    if (lane + u * 32 < n4) {
      cur.a[u] = st[(u * 2) * 32];
      cur.b[u] = st[(u * 2 + 1) * 32];
    } else {
      cur.a[u] = f4_zero();
      cur.b[u] = f4_zero();
    }
  }
}

// One GEMV phase.  `cur` holds this warp's first tile (mega_prefetch); on return it holds
// the first tile of `next_ph` (if next_ph.W != nullptr).
template <int PRO, int EPI>
__device__ __noinline__ void mega_gemv(const MegaParams& p, const int l, const int pos, const MegaPhase ph,
                                          const MegaPhase next_ph, const float* vin, const float* rms_w,
                                          float* write_x, unsigned char* xs, double* red_scratch, float4* stage,
                                          const uint64_t pol, float& bv, int& bi) {
  typedef XVec<true> XV;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = ph.n, n4 = n >> 2;
  const int npairs = ph.rows >> 1;
  const int pair0 = (int)(((long long)npairs * blockIdx.x) / gridDim.x);
  const int pair1 = (int)(((long long)npairs * (blockIdx.x + 1)) / gridDim.x);
  const int tpp = (n4 + 32 * kMU - 1) / (32 * kMU);
  const int my_first = pair0 + warp;
  const int my_pairs = my_first < pair1 ? (pair1 - my_first + kMegaWarps - 1) / kMegaWarps : 0;
  const float4* W4 = reinterpret_cast<const float4*>(ph.W);

  // prologue: activation vector -> shared memory (as doubles)
  {
    const float4* src4 = reinterpret_cast<const float4*>(vin);
    double ss = 0.0;
    for (int j = threadIdx.x; j < n4; j += kMegaThreads) {
      const float4 v = ld_cg4(src4 + j);
      if (PRO == PRO_COPY) {
        XV::store(xs, n4, j, v);
      } else {
        ss += (double)v.x * (double)v.x + (double)v.y * (double)v.y + (double)v.z * (double)v.z +
              (double)v.w * (double)v.w;
      }
      if (write_x != nullptr && blockIdx.x == 0) reinterpret_cast<float4*>(write_x)[j] = v;
    }
    if (PRO == PRO_RMS) {
      ss = warp_sum_f64(ss);
      if (lane == 0) red_scratch[warp] = ss;
      __syncthreads();
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < kMegaWarps; ++w) tot += red_scratch[w];
      tot /= (double)n;
      tot = 1.0 / sqrt(1e-5 + tot);
      const float4* rw4 = reinterpret_cast<const float4*>(rms_w);
      for (int j = threadIdx.x; j < n4; j += kMegaThreads) {
        const float4 v = ld_cg4(src4 + j);
        const float4 w = __ldg(rw4 + j);
        float4 o;
        o.x = (float)((double)w.x * (tot * (double)v.x));
        o.y = (float)((double)w.y * (tot * (double)v.y));
        o.z = (float)((double)w.z * (tot * (double)v.z));
        o.w = (float)((double)w.w * (tot * (double)v.w));
        XV::store(xs, n4, j, o);
      }
    }
  }
  __syncthreads();

  MegaTile cur, nxt;
  int pair = my_first, jt = 0;
  double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
  const int total = my_pairs * tpp;
  mega_take_prefetched(cur, stage, n4);
  for (int t = 0; t < total; ++t) {
    int npair = pair, njt = jt + 1;
    if (njt == tpp) {
      njt = 0;
      npair = pair + kMegaWarps;
    }
    if (t + 1 < total) {
      const float4* w0 = W4 + (size_t)(2 * npair) * n4;
      mega_load_tile(nxt, w0, w0 + n4, njt * 32 * kMU + lane, n4, pol);
    } else if (next_ph.W != nullptr) {
      mega_prefetch(next_ph, stage);  // keep the HBM pipe busy through the barrier
    }
    {
      const int j0 = jt * 32 * kMU + lane;
#pragma unroll
      for (int u = 0; u < kMU; ++u) {
        const int idx = j0 + u * 32;
        if (idx < n4) {
          double xv[4];
          XV::load(xs, n4, idx, xv);
          double r0 = acc[0][u & 1], r1 = acc[1][u & 1];
          r0 = fma((double)cur.a[u].x, xv[0], r0);
          r1 = fma((double)cur.b[u].x, xv[0], r1);
          r0 = fma((double)cur.a[u].y, xv[1], r0);
          r1 = fma((double)cur.b[u].y, xv[1], r1);
          r0 = fma((double)cur.a[u].z, xv[2], r0);
          r1 = fma((double)cur.b[u].z, xv[2], r1);
          r0 = fma((double)cur.a[u].w, xv[3], r0);
          r1 = fma((double)cur.b[u].w, xv[3], r1);
          acc[0][u & 1] = r0;
          acc[1][u & 1] = r1;
        }
      }
    }
    if (jt == tpp - 1) {
      double d0 = warp_sum_f64(acc[0][0] + acc[0][1]);
      double d1 = warp_sum_f64(acc[1][0] + acc[1][1]);
      acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0.0;
      if (lane == 0) {
        const float s0 = (float)d0, s1 = (float)d1;
        const int r = 2 * pair;
        if (EPI == EPI_QKV) {
          const int seg = r / p.D, i = r - seg * p.D;
          const int h = i / p.hs, c = i - h * p.hs;
          const size_t row = (size_t)l * p.kv_layer + ((size_t)h * p.steps + pos) * p.hs + c;
          if (seg == 2) {
            p.vc[row] = s0;
            p.vc[row + 1] = s1;
          } else {
            const double fr = (double)__ldg(p.fcr + (size_t)pos * (p.hs / 2) + c / 2);
            const double fi = (double)__ldg(p.fci + (size_t)pos * (p.hs / 2) + c / 2);
            const float o0 = (float)((double)s0 * fr - (double)s1 * fi);
            const float o1 = (float)((double)s0 * fi + (double)s1 * fr);
            float* dst = seg == 0 ? p.q + i : p.kc + row;
            dst[0] = o0;
            dst[1] = o1;
          }
        } else if (EPI == EPI_RESID) {
          p.x[r] = (float)((double)ld_cg(p.x + r) + (double)s0);
          p.x[r + 1] = (float)((double)ld_cg(p.x + r + 1) + (double)s1);
        } else if (EPI == EPI_SWIGLU) {
          const double hv = (double)s0;
          const float silu = (float)(hv * (1.0 / (1.0 + exp(-hv))));
          p.hb[pair] = (float)((double)silu * (double)s1);
        } else {
          p.logits[r] = s0;
          p.logits[r + 1] = s1;
          argmax_consider(s0, r, bv, bi);
          argmax_consider(s1, r + 1, bv, bi);
        }
      }
    }
    pair = npair;
    jt = njt;
    cur = nxt;
  }
  if (total == 0 && next_ph.W != nullptr) mega_prefetch(next_ph, stage);
}

// attention for head h of one sequence, whole CTA (llama2.ts:244-267); exact two-pass softmax
__device__ __noinline__ void mega_attention(const MegaParams& p, int l, int h, int pos, float* sc,
                                               float* s_red /*[kMegaWarps][kAttnMaxHs]*/, float* s_wmax,
                                               double* s_wsum) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hs = p.hs, hs4 = hs >> 2;
  const int G = hs4 <= 16 ? 16 : 32;
  const int subs = 32 / G;
  const int sub = lane / G, c4 = lane % G;
  const float* kbase = p.kc + (size_t)l * p.kv_layer + (size_t)h * p.steps * hs;
  const float* vbase = p.vc + (size_t)l * p.kv_layer + (size_t)h * p.steps * hs;
  float4 qv[2];
  {
    const float4* q4 = reinterpret_cast<const float4*>(p.q + (size_t)h * hs);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = c4 + i * G;
      qv[i] = c < hs4 ? ld_cg4(q4 + c) : f4_zero();
    }
  }
  const double sqrt_hs = sqrt((double)hs);
  const int n_t = pos + 1;
  const int rows_per_pass = kMegaWarps * subs;
  float lmax = -INFINITY;
  for (int base = warp * subs; base < n_t; base += rows_per_pass) {
    const int t = base + sub;
    const bool valid = t < n_t;
    double d = 0.0;
    if (valid) {
      const float4* k4 = reinterpret_cast<const float4*>(kbase + (size_t)t * hs);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int c = c4 + k * G;
        if (c < hs4) {
          const float4 kv = ld_cg4(k4 + c);
          d = fma((double)qv[k].x, (double)kv.x, d);
          d = fma((double)qv[k].y, (double)kv.y, d);
          d = fma((double)qv[k].z, (double)kv.z, d);
          d = fma((double)qv[k].w, (double)kv.w, d);
        }
      }
    }
    for (int o = G >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (valid && c4 == 0) {
      const float sv = (float)(d / sqrt_hs);
      sc[t] = sv;
      lmax = fmaxf(lmax, sv);
    }
  }
  lmax = warp_max_f32(lmax);
  if (lane == 0) s_wmax[warp] = lmax;
  __syncthreads();
  float gmax = s_wmax[0];
#pragma unroll
  for (int w = 1; w < kMegaWarps; ++w) gmax = fmaxf(gmax, s_wmax[w]);
  double lsum = 0.0;
  for (int t = tid; t < n_t; t += kMegaThreads) {
    const float ex = (float)exp((double)sc[t] - (double)gmax);
    sc[t] = ex;
    lsum += (double)ex;
  }
  lsum = warp_sum_f64(lsum);
  if (lane == 0) s_wsum[warp] = lsum;
  __syncthreads();
  double gsum = 0.0;
#pragma unroll
  for (int w = 0; w < kMegaWarps; ++w) gsum += s_wsum[w];
  for (int t = tid; t < n_t; t += kMegaThreads) sc[t] = (float)((double)sc[t] / gsum);
  __syncthreads();

  float4 acc[2] = {f4_zero(), f4_zero()};
  for (int t = warp * subs + sub; t < n_t; t += rows_per_pass) {
    const float a = sc[t];
    const float4* v4 = reinterpret_cast<const float4*>(vbase + (size_t)t * hs);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = c4 + k * G;
      if (c < hs4) {
        const float4 vv = ld_cg4(v4 + c);
        acc[k].x = fmaf(a, vv.x, acc[k].x);
        acc[k].y = fmaf(a, vv.y, acc[k].y);
        acc[k].z = fmaf(a, vv.z, acc[k].z);
        acc[k].w = fmaf(a, vv.w, acc[k].w);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    for (int o = 16; o >= G; o >>= 1) {
      acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, o);
      acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, o);
      acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, o);
      acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, o);
    }
    const int c = c4 + k * G;
    if (sub == 0 && c < hs4) reinterpret_cast<float4*>(s_red + (size_t)warp * kAttnMaxHs)[c] = acc[k];
  }
  __syncthreads();
  if (tid < hs) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kMegaWarps; ++w) s += s_red[(size_t)w * kAttnMaxHs + tid];
    p.xb[(size_t)h * hs + tid] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kMegaThreads, 1) mega_decode_kernel(const __grid_constant__ MegaParams p) {
  extern __shared__ __align__(16) unsigned char mega_smem[];
  __shared__ double red_scratch[kMegaWarps];
  __shared__ float s_wmax[kMegaWarps];
  __shared__ double s_wsum[kMegaWarps];
  __shared__ float s_bv[kMegaWarps];
  __shared__ int s_bi[kMegaWarps];
  // dynamic region: [0, 64 KB) prefetch stage | activation vector; the attention phase
  // reuses the activation part: scores | per-warp partial outputs
  float4* stage = reinterpret_cast<float4*>(mega_smem);
  unsigned char* xs = mega_smem + (size_t)kMegaWarps * kMU * 2 * 32 * 16;
  float* sc = reinterpret_cast<float*>(xs);
  float* s_red = sc + ((p.steps + 3) & ~3);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = p.D, F = p.F, L = p.L;
  const uint64_t pol = make_l2_policy(p.evict_first != 0);
  MegaPhase ph_qkv0 = {p.wqkv, 3 * D, D};
  mega_prefetch(ph_qkv0, stage);

#pragma unroll 1
  for (int step = 0; step < p.n_steps; ++step) {
    const int token = ld_cg_i32(p.ctl + CTL_HDR);
    const int pos = ld_cg_i32(p.ctl + CTL_HDR + 1);
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const MegaPhase ph_qkv = {p.wqkv + (size_t)l * 3 * D * D, 3 * D, D};
      const MegaPhase ph_wo = {p.wo + (size_t)l * D * D, D, D};
      const MegaPhase ph_w13 = {p.w13 + (size_t)l * 2 * F * D, 2 * F, D};
      const MegaPhase ph_w2 = {p.w2 + (size_t)l * D * F, D, F};
      const MegaPhase ph_next = (l + 1 < L) ? MegaPhase{p.wqkv + (size_t)(l + 1) * 3 * D * D, 3 * D, D}
                                            : MegaPhase{p.wcls, p.V, D};
      // rmsnorm -> q,k,v -> RoPE -> KV write (layer 0: x := embedding row)
      mega_gemv<PRO_RMS, EPI_QKV>(p, l, pos, ph_qkv, ph_wo, l == 0 ? p.tok_emb + (size_t)token * D : p.x,
                                  p.rms_att + (size_t)l * D, l == 0 ? p.x : nullptr, xs, red_scratch, stage,
                                  pol, bv, bi);
      grid_sync(p.bar);
      for (int h = blockIdx.x; h < p.H; h += gridDim.x) mega_attention(p, l, h, pos, sc, s_red, s_wmax, s_wsum);
      grid_sync(p.bar);
      mega_gemv<PRO_COPY, EPI_RESID>(p, l, pos, ph_wo, ph_w13, p.xb, nullptr, nullptr, xs, red_scratch, stage,
                                     pol, bv, bi);
      grid_sync(p.bar);
      mega_gemv<PRO_RMS, EPI_SWIGLU>(p, l, pos, ph_w13, ph_w2, p.x, p.rms_ffn + (size_t)l * D, nullptr, xs,
                                     red_scratch, stage, pol, bv, bi);
      grid_sync(p.bar);
      mega_gemv<PRO_COPY, EPI_RESID>(p, l, pos, ph_w2, ph_next, p.hb, nullptr, nullptr, xs, red_scratch, stage,
                                     pol, bv, bi);
      grid_sync(p.bar);
    }
    // final rmsnorm -> classifier -> argmax; the next step starts with layer 0's q/k/v rows
    const MegaPhase ph_cls = {p.wcls, p.V, D};
    const MegaPhase ph_first = (step + 1 < p.n_steps) ? ph_qkv0 : MegaPhase{nullptr, 0, 0};
    mega_gemv<PRO_RMS, EPI_LOGITS>(p, 0, pos, ph_cls, ph_first, p.x, p.rms_final, nullptr, xs, red_scratch,
                                   stage, pol, bv, bi);
    if (lane == 0) {
      s_bv[warp] = bv;
      s_bi[warp] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kMegaWarps; ++w) argmax_consider(s_bv[w], s_bi[w], bv, bi);
      p.blk_val[blockIdx.x] = bv;
      p.blk_idx[blockIdx.x] = bi;
    }
    grid_sync(p.bar);
    if (blockIdx.x == 0 && warp == 0) {
      float v = -INFINITY;
      int i = 0x7fffffff;
      for (int g = lane; g < (int)gridDim.x; g += 32) argmax_consider(ld_cg(p.blk_val + g), ld_cg_i32(p.blk_idx + g), v, i);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        argmax_consider(ov, oi, v, i);
      }
      if (lane == 0) {
        const float l0 = ld_cg(p.logits);
        if (i == 0x7fffffff || l0 != l0) i = 0;
        const int st = ld_cg_i32(p.ctl + CTL_STEP);
        int chosen = i;
        if (ld_cg_i32(p.ctl + CTL_USE_FORCED)) {
          const int f = p.forced[st];
          if (f >= 0) chosen = f;
        }
        p.next[0] = i;
        p.out_tokens[st] = chosen;
        if (ld_cg_i32(p.ctl + CTL_ADVANCE)) {
          p.ctl[CTL_HDR] = chosen;
          p.ctl[CTL_HDR + 1] = pos + 1;
          p.ctl[CTL_STEP] = st + 1;
        }
      }
    }
    if (step + 1 < p.n_steps) grid_sync(p.bar);
  }
}

}  // namespace l2b
