// fused_qkv_attn.cuh -- batch-1: rmsnorm + q/k/v rows + RoPE + KV-cache write AND the attention
// of the same layer (llama2.ts:216-267) in one kernel.
//
// Why: attention only depends on the q/k/v rows of ITS head.  One thread-block cluster per
// head computes that head's 3*head_size rows (split over the cluster's CTAs, same row-pair
// GEMV as decode_kernels.cuh), hands q and the new k/v row to every CTA of the cluster through
// distributed shared memory, and after one cluster barrier runs the attention of that head
// split over time (exact two-pass softmax through DSMEM, as l2b_attn_decode_kernel).  That removes
// one of the five kernel boundaries of a layer (~6 us each on Llama-2-7B, 32 per token) and
// the attention kernel's own start-up; the K/V rows of earlier positions are requested with
// bulk async copies right after griddepcontrol.wait, so they land while the GEMV part runs.
//
// grid (CS, n_heads), cluster (CS,1,1), 512 threads.  The newest K/V row never travels through
// global memory inside the kernel (it is also written to the cache for later steps).
#pragma once
#include <math.h>

#include "common.cuh"
#include "decode_kernels.cuh"

namespace l2b {

constexpr int kFThreads = 512;
constexpr int kFWarps = kFThreads / 32;

struct QkvAttnParams {
  const float* W;        // wqkv of this layer, [3D][D]
  int D, H, hs, steps;
  const float* vin;      // x (residual stream)
  const float* rms_w;
  const float* tok_emb;  // layer 0: x := embedding row
  const int* tokp;
  const int* posp;
  float* x;
  float* q;              // [D], kept for state read-back
  float* kc;             // this layer's caches, sequence 0: [H][steps][hs]
  float* vc;
  const float* fcr;
  const float* fci;
  float* xb;             // attention output [D]
  int tileT, sc_cap, stage_bytes;
  int evict_first;
  int l2_prefetch;
  // the attention part leaves HBM idle: meanwhile pull the NEXT kernel's weights (wo of this
  // layer) into L2.  The K/V copies of this CTA were issued long before, nothing queues behind.
  const unsigned char* pf_ptr;
  long long pf_bytes;
  // software hand-over (common.cuh)
  const int* sync_wait;
  int sync_target;
  int* sync_done;
  // tensor parallel (TP = true instantiation): this rank owns H of the model's heads; wqkv holds
  // its rows as three segments of seg_rows; x arrives as an LL replica tagged with sequence
  // epoch + 1 + tp_wait_idx (layer 0: the embedding row); the output slice goes to every peer's
  // xb replica as LL words tagged epoch + 1 + tp_out_idx
  int seg_rows;            // rows per q/k/v segment of W (== D when not sharded)
  int tp_size;
  int tp_ll_in;
  const int* tp_epoch;
  int tp_wait_idx, tp_out_idx;
  int xb_off;              // column of head 0 of this rank in the gathered xb
  int* tp_err;
  float* peer_xb[kMaxTp];
};

__device__ __forceinline__ void dsmem_st_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

template <bool TP>
__device__ __forceinline__ void qkv_attn_body(const QkvAttnParams& p) {
  typedef XVec<true> XV;
  extern __shared__ __align__(128) unsigned char f_smem[];
  // dynamic: activation vector (doubles) | K/V ring | scores
  __shared__ __align__(8) uint64_t full_bar[kAttnStages];
  __shared__ __align__(8) uint64_t empty_bar[kAttnStages];
  __shared__ double red_scratch[kFWarps];
  __shared__ __align__(16) float s_q[kAttnMaxHs];   // written by every CTA of the cluster (DSMEM)
  __shared__ __align__(16) float s_k[kAttnMaxHs];
  __shared__ __align__(16) float s_v[kAttnMaxHs];
  __shared__ float s_red[kFWarps][kAttnMaxHs];
  __shared__ float s_wmax[kFWarps];
  __shared__ double s_wsum[kFWarps];
  __shared__ float c_out[kAttnMaxHs];
  __shared__ float c_max;
  __shared__ double c_sum;

  griddep_launch_dependents();
  // "This CTA has started": a peer's shared memory may only be written once that block runs (found by
  // compute-sanitizer racecheck).  Split barrier: arrive here, wait (whole warp, once) right before the
  // warp's first remote store / before the first full cluster barrier -- by then every peer has long
  // arrived; an aligned wait right after the prologue cost 0.9 us per launch on stories15M (the CTAs
  // of a cluster do not start at the same instant).
  cluster_arrive();
  bool entered_wait_done = false;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank(), CS = cluster_nctarank();
  const int h = blockIdx.y;
  const int D = p.D, n4 = D >> 2, hs = p.hs, hs4 = hs >> 2;
  unsigned char* xs = f_smem;
  float* ring = reinterpret_cast<float*>(f_smem + (size_t)D * 8);
  float* sc = ring + (size_t)kAttnStages * (p.stage_bytes / 4);

  if (tid == 0) {
    for (int s = 0; s < kAttnStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kFWarps);
    }
    mbar_fence_init();
  }

  // ---- GEMV part: this CTA's share of the head's 3*hs rows, as row pairs ----
  const int half = hs >> 1;                 // row pairs per segment (q, k, v)
  const int P = 3 * half;
  const int pi0 = (P * (int)rank) / (int)CS, pi1 = (P * ((int)rank + 1)) / (int)CS;
  const float4* W4 = reinterpret_cast<const float4*>(p.W);
  const uint64_t pol = make_l2_policy(p.evict_first != 0);
  const int tpp = (n4 + 32 * kU - 1) / (32 * kU);
  auto first_row = [&](int pi) {            // global row of pair pi in wqkv
    const int seg = pi / half;
    return seg * p.seg_rows + h * hs + 2 * (pi - seg * half);
  };
  PairTile cur, nxt;
  int pi = pi0 + warp;
  if (pi < pi1) {
    const float4* w0 = W4 + (size_t)first_row(pi) * n4;
    load_pair_tile(cur, w0, w0 + n4, lane, n4, pol);
  }
  if (p.l2_prefetch > 0 && lane == 0 && pi + kFWarps < pi1) {  // head of this warp's second pair -> L2
    size_t want = (size_t)p.l2_prefetch / kFWarps;
    const size_t pair_bytes = (size_t)2 * D * sizeof(float);
    if (want > pair_bytes) want = pair_bytes;
    want &= ~(size_t)15;
    if (want > 0)
      prefetch_l2_bulk(reinterpret_cast<const unsigned char*>(p.W + (size_t)first_row(pi + kFWarps) * D), (uint32_t)want);
  }
  __syncthreads();  // mbarrier init visible to the CTA

  // ---- everything below may depend on the previous kernel ----
  if (p.sync_wait != nullptr) {
    soft_wait(p.sync_wait, p.sync_target);
  } else {
    griddep_wait();
  }
  const int pos = ld_act_i32(p.posp);
  const int n_t = pos + 1;
  int tp_seq = 0;
  if (TP) tp_seq = ld_act_i32(p.tp_epoch) + 1;

  // attention chunk of this CTA; rows before `pos` come from the cache (earlier steps), the
  // newest row from shared memory
  const int chunk = (n_t + (int)CS - 1) / (int)CS;
  const int t0 = (int)rank * chunk;
  int nT = n_t - t0;
  nT = nT < 0 ? 0 : (nT > chunk ? chunk : nT);
  const bool owns_cur = nT > 0 && t0 + nT == n_t;
  const int nC = owns_cur ? nT - 1 : nT;          // cached rows of this chunk
  const int tileT = p.tileT;
  const int nTiles = (nC + tileT - 1) / tileT;
  const int total = 2 * nTiles;
  const size_t head_off = ((size_t)h * p.steps) * hs;
  const float* kbase = p.kc + head_off + (size_t)t0 * hs;
  const float* vbase = p.vc + head_off + (size_t)t0 * hs;
  const int stage_floats = p.stage_bytes / 4;
  auto issue = [&](int j) {
    const int s = j % kAttnStages;
    if (j >= kAttnStages) mbar_wait(&empty_bar[s], ((j / kAttnStages) - 1) & 1);
    const int tile = j < nTiles ? j : j - nTiles;
    const float* base = j < nTiles ? kbase : vbase;
    const int tt0 = tile * tileT;
    const int cnt = min(tileT, nC - tt0);
    const uint32_t bytes = (uint32_t)cnt * hs * 4u;
    mbar_arrive_expect_tx(&full_bar[s], bytes);
    bulk_g2s(ring + (size_t)s * stage_floats, base + (size_t)tt0 * hs, bytes, &full_bar[s]);
  };
  if (tid == 0)
    for (int j = 0; j < kAttnStages - 1 && j < total; ++j) issue(j);

  // prologue: x (or the embedding row) -> rmsnorm -> shared memory, llama2.ts:211, 216
  {
    const float* src = p.tok_emb != nullptr ? p.tok_emb + (size_t)ld_act_i32(p.tokp) * D : p.vin;
    const float4* src4 = reinterpret_cast<const float4*>(src);
    const bool write_x = p.tok_emb != nullptr && rank == 0 && h == 0;
    const bool ll_in = TP && p.tp_ll_in != 0;
    double ss = 0.0;
    for (int j = tid; j < n4; j += kFThreads) {
      const float4 v = ll_in ? ll_load4(p.vin, j, true, tp_seq + p.tp_wait_idx, p.tp_err) : ld_act4(src4 + j);
      ss += (double)v.x * (double)v.x + (double)v.y * (double)v.y + (double)v.z * (double)v.z +
            (double)v.w * (double)v.w;
      if (write_x) {
        if (TP) {  // x is an LL replica on every rank
          uint4* xl = reinterpret_cast<uint4*>(p.x) + 2 * (size_t)j;
          xl[0] = make_uint4(__float_as_uint(v.x), 0u, __float_as_uint(v.y), 0u);
          xl[1] = make_uint4(__float_as_uint(v.z), 0u, __float_as_uint(v.w), 0u);
        } else {
          reinterpret_cast<float4*>(p.x)[j] = v;
        }
      }
    }
    ss = warp_sum_f64(ss);
    if (lane == 0) red_scratch[warp] = ss;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < kFWarps; ++w) tot += red_scratch[w];
    tot /= (double)D;
    tot = 1.0 / sqrt(1e-5 + tot);
    const float4* rw4 = reinterpret_cast<const float4*>(p.rms_w);
    for (int j = tid; j < n4; j += kFThreads) {
      const float4 v = ll_in ? ll_load4(p.vin, j, false, 0, nullptr) : ld_act4(src4 + j);
      const float4 w = __ldg(rw4 + j);
      float4 o;
      o.x = (float)((double)w.x * (tot * (double)v.x));
      o.y = (float)((double)w.y * (tot * (double)v.y));
      o.z = (float)((double)w.z * (tot * (double)v.z));
      o.w = (float)((double)w.w * (tot * (double)v.w));
      XV::store(xs, n4, j, o);
    }
  }
  __syncthreads();

  {
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    double tot0 = 0.0, tot1 = 0.0;
    int jt = 0;
    bool have = pi < pi1;
    while (have) {
      int npi = pi, njt = jt + 1;
      if (njt == tpp) {
        njt = 0;
        npi = pi + kFWarps;
      }
      const bool more = npi < pi1;
      if (more) {
        const float4* w0 = W4 + (size_t)first_row(npi) * n4;
        load_pair_tile(nxt, w0, w0 + n4, njt * 32 * kU + lane, n4, pol);
      }
      {
        const int j0 = jt * 32 * kU + lane;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
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
      if ((jt + 1) % kChunkTiles == 0 || jt == tpp - 1) {   // a K-chunk is complete: the canonical summation
        tot0 += acc[0][0] + acc[0][1];                      // order of l2b_rowpair_matvec_kernel (lane totals in
        tot1 += acc[1][0] + acc[1][1];                      // chunk order, then one shuffle tree): same bits
        acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0.0;
      }
      if (jt == tpp - 1) {
        if (!entered_wait_done) {   // warp-uniform: the whole warp completes the entry phase before its first remote store
          cluster_wait();
          entered_wait_done = true;
        }
        const double d0 = warp_sum_f64(tot0), d1 = warp_sum_f64(tot1);
        tot0 = tot1 = 0.0;
        if (lane == 0) {
          const float s0 = (float)d0, s1 = (float)d1;
          const int seg = pi / half, c = 2 * (pi - seg * half);
          const size_t row = head_off + (size_t)pos * hs + c;
          float o0 = s0, o1 = s1;
          float* sm = s_v;
          if (seg == 2) {  // value row pair, llama2.ts:240
            p.vc[row] = s0;
            p.vc[row + 1] = s1;
          } else {         // RoPE, llama2.ts:223-236
            const double fr = (double)__ldg(p.fcr + (size_t)pos * half + c / 2);
            const double fi = (double)__ldg(p.fci + (size_t)pos * half + c / 2);
            o0 = (float)((double)s0 * fr - (double)s1 * fi);
            o1 = (float)((double)s0 * fi + (double)s1 * fr);
            if (seg == 0) {
              sm = s_q;
              p.q[h * hs + c] = o0;
              p.q[h * hs + c + 1] = o1;
            } else {
              sm = s_k;
              p.kc[row] = o0;
              p.kc[row + 1] = o1;
            }
          }
          for (uint32_t r = 0; r < CS; ++r) {
            const uint32_t a = dsmem_addr(sm + c, r);
            dsmem_st_f32(a, o0);
            dsmem_st_f32(a + 4, o1);
          }
        }
      }
      have = more;
      pi = npi;
      jt = njt;
      cur = nxt;
    }
  }
  if (p.pf_bytes > 0 && lane == 0) {
    const long long n_warp = (long long)gridDim.x * gridDim.y * kFWarps;
    const long long me = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * kFWarps + warp;
    const long long per = ((p.pf_bytes / n_warp) + 15) & ~15LL;
    long long off = me * per;
    const long long end = off + per < p.pf_bytes ? off + per : p.pf_bytes;
    for (; off + 16 <= end; off += 32768) {
      const long long len = (end - off < 32768 ? end - off : 32768) & ~15LL;
      if (len > 0) prefetch_l2_bulk(p.pf_ptr + off, (uint32_t)len);
    }
  }
  if (!entered_wait_done) cluster_wait();   // warps without a row pair complete the entry phase here
  cluster_sync_all();  // q, k, v of this head are in every CTA's shared memory

  // ---- attention (llama2.ts:244-267) over rows [t0, t0 + nT) ----
  const int G = hs4 <= 16 ? 16 : 32;
  const int subs = 32 / G;
  const int sub = lane / G, c4 = lane % G;
  float4 qv[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = c4 + i * G;
    qv[i] = c < hs4 ? reinterpret_cast<const float4*>(s_q)[c] : f4_zero();
  }
  const double sqrt_hs = sqrt((double)hs);
  const int rows_per_pass = kFWarps * subs;

  float lmax = -INFINITY;
  for (int i = 0; i < nTiles; ++i) {
    if (tid == 0 && i + kAttnStages - 1 < total) issue(i + kAttnStages - 1);
    __syncwarp();
    const int s = i % kAttnStages;
    mbar_wait(&full_bar[s], (i / kAttnStages) & 1);
    const float4* st4 = reinterpret_cast<const float4*>(ring + (size_t)s * stage_floats);
    const int tt0 = i * tileT;
    const int cnt = min(tileT, nC - tt0);
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
        const float sv = (float)(d / sqrt_hs);
        sc[tt0 + tt] = sv;
        lmax = fmaxf(lmax, sv);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
  if (owns_cur && warp == 0) {  // the newest row, from shared memory
    double d = 0.0;
    if (sub == 0) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int c = c4 + k * G;
        if (c < hs4) {
          const float4 kv = reinterpret_cast<const float4*>(s_k)[c];
          d = fma((double)qv[k].x, (double)kv.x, d);
          d = fma((double)qv[k].y, (double)kv.y, d);
          d = fma((double)qv[k].z, (double)kv.z, d);
          d = fma((double)qv[k].w, (double)kv.w, d);
        }
      }
    }
    for (int o = G >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) {
      const float sv = (float)(d / sqrt_hs);
      sc[nT - 1] = sv;
      lmax = fmaxf(lmax, sv);
    }
  }

  lmax = warp_max_f32(lmax);
  if (lane == 0) s_wmax[warp] = lmax;
  __syncthreads();
  if (tid == 0) {
    float m = s_wmax[0];
    for (int w = 1; w < kFWarps; ++w) m = fmaxf(m, s_wmax[w]);
    c_max = m;
  }
  cluster_sync_all();
  float gmax = -INFINITY;
  for (uint32_t r = 0; r < CS; ++r) gmax = fmaxf(gmax, dsmem_ld_f32(dsmem_addr(&c_max, r)));
  double lsum = 0.0;
  for (int t = tid; t < nT; t += kFThreads) {
    const float e = (float)exp((double)sc[t] - (double)gmax);
    sc[t] = e;
    lsum += (double)e;
  }
  lsum = warp_sum_f64(lsum);
  if (lane == 0) s_wsum[warp] = lsum;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kFWarps; ++w) s += s_wsum[w];
    c_sum = s;
  }
  cluster_sync_all();
  double gsum = 0.0;
  for (uint32_t r = 0; r < CS; ++r) gsum += dsmem_ld_f64(dsmem_addr(&c_sum, r));
  for (int t = tid; t < nT; t += kFThreads) sc[t] = (float)((double)sc[t] / gsum);
  __syncthreads();

  float4 acc[2] = {f4_zero(), f4_zero()};
  for (int i = nTiles; i < total; ++i) {
    if (tid == 0 && i + kAttnStages - 1 < total) issue(i + kAttnStages - 1);
    __syncwarp();
    const int s = i % kAttnStages;
    mbar_wait(&full_bar[s], (i / kAttnStages) & 1);
    const float4* st4 = reinterpret_cast<const float4*>(ring + (size_t)s * stage_floats);
    const int tt0 = (i - nTiles) * tileT;
    const int cnt = min(tileT, nC - tt0);
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
  if (owns_cur && warp == 0 && sub == 0) {
    const float a = sc[nT - 1];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = c4 + k * G;
      if (c < hs4) {
        const float4 vv = reinterpret_cast<const float4*>(s_v)[c];
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
    if (sub == 0 && c < hs4) reinterpret_cast<float4*>(&s_red[warp][0])[c] = acc[k];
  }
  __syncthreads();
  if (tid < hs) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kFWarps; ++w) s += s_red[w][tid];
    c_out[tid] = s;
  }
  cluster_sync_all();
  if (rank == 0 && tid < hs) {
    float s = 0.f;
    for (uint32_t r = 0; r < CS; ++r) s += dsmem_ld_f32(dsmem_addr(&c_out[tid], r));
    if (!TP) {
      p.xb[(size_t)h * hs + tid] = s;
    } else {
      const size_t o = (size_t)p.xb_off + (size_t)h * hs + tid;
      const uint32_t sq = (uint32_t)(tp_seq + p.tp_out_idx);
      for (int r = 0; r < p.tp_size; ++r)
        st_sys_u2(reinterpret_cast<uint2*>(p.peer_xb[r]) + o, __float_as_uint(s), sq);
    }
  }
  cluster_sync_all();  // keep every CTA's shared memory alive until rank 0 has read it
  if (p.sync_done != nullptr) soft_signal(p.sync_done);
}
__global__ void __launch_bounds__(kFThreads, 1) l2b_qkv_attn_kernel(const __grid_constant__ QkvAttnParams p) {
  qkv_attn_body<false>(p);
}
__global__ void __launch_bounds__(kFThreads, 1) l2b_qkv_attn_tp_kernel(const __grid_constant__ QkvAttnParams p) {
  qkv_attn_body<true>(p);
}

}  // namespace l2b
