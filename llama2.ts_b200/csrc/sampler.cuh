// sampler.cuh -- SURVEY.md section 8f rank 1: temperature, softmax and the samplers of
// llama2.ts:476-494 / :368-394 on the device, so that a sampled token costs a 4-byte
// read-back instead of 128 KB of logits plus a 32000-element sort on the host.
//
// The host keeps the xorshift RNG: it draws its ONE random_f32() per sampled token
// (llama2.ts:370 / :388) exactly as before and passes the value in.
//
// One CTA of 1024 threads (vocab is 32000: 128 KB, L2-resident); thread t owns the contiguous
// index range [t*c, (t+1)*c) so that prefix sums run in index order.  Rounding points follow
// the reference: logits/temperature and exp() are stored as f32, sums are f64.  The only
// deviation is the ASSOCIATION of the f64 sums (per-thread sequential, then a block scan),
// 1e-16 relative, visible only if the random number falls within that distance of a CDF step.
// Top-p: candidates p >= (1-topp)/(V-1) (everything smaller cannot be reached by the
// reference's walk over the sorted array), compacted in index order, ranked by (prob desc,
// index asc) -- the order of the reference's stable sort; a rank count for up to 2048
// candidates, a stable 4-bit LSD radix sort in shared-memory counters above that (a flat
// distribution makes all 32000 entries candidates) -- and then walked SEQUENTIALLY by one
// thread exactly like llama2.ts:382-393, including the exclusive `i < lastIdx` quirk and the
// fallback to token 0.  If the walk would leave the candidate set it is redone over all V.
#pragma once
#include <math.h>

#include "common.cuh"

namespace l2b {

constexpr int kSampThreads = 1024;

struct SampleParams {
  const float* logits;  // [V]
  int V;
  double temperature, topp;
  double rand01;        // the host's random_f32(), widened
  float* probs;         // [V] scratch
  float* cand_p;        // [V]
  int* cand_i;          // [V]
  float* sort_p;        // [V]
  int* sort_i;          // [V]
  int* next;            // result
};

__device__ __forceinline__ double block_excl_scan_f64(double v, double* s_warp, double* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    double w = s_warp[lane];
    double winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    s_warp[lane] = winc - w;          // exclusive prefix of the warp sums
    if (lane == 31) s_warp[32] = winc;  // grand total
  }
  __syncthreads();
  const double res = s_warp[warp] + (inc - v);
  *total = s_warp[32];
  __syncthreads();
  return res;
}

__device__ __forceinline__ int block_excl_scan_i32(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    s_warp[lane] = winc - w;
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  const int res = s_warp[warp] + (inc - v);
  *total = s_warp[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kSampThreads) l2b_sample_kernel(const __grid_constant__ SampleParams p) {
  __shared__ double s_d[33];
  __shared__ int s_i[33];
  __shared__ float s_f[32];
  __shared__ float s_tile_p[2048];
  __shared__ int s_tile_i[2048];
  __shared__ int s_retry, s_k;
  extern __shared__ int s_cnt[];  // [16][kSampThreads] radix counters
  griddep_launch_dependents();
  griddep_wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = p.V;
  const int c = (V + kSampThreads - 1) / kSampThreads;
  const int i0 = tid * c < V ? tid * c : V, i1 = (tid + 1) * c < V ? (tid + 1) * c : V;

  // logits[q] /= temperature (llama2.ts:481-483), then softmax (:181-194)
  float tmax = -INFINITY;
  for (int i = i0; i < i1; ++i) {
    const float x = (float)((double)ld_act(p.logits + i) / p.temperature);
    p.probs[i] = x;
    tmax = fmaxf(tmax, x);
  }
  tmax = warp_max_f32(tmax);
  if (lane == 0) s_f[warp] = tmax;
  __syncthreads();
  float gmax = s_f[0];
  for (int w = 1; w < 32; ++w) gmax = fmaxf(gmax, s_f[w]);
  double lsum = 0.0;
  for (int i = i0; i < i1; ++i) {
    const float e = (float)exp((double)p.probs[i] - (double)gmax);
    p.probs[i] = e;
    lsum += (double)e;
  }
  double total;
  block_excl_scan_f64(lsum, s_d, &total);
  double psum = 0.0;
  for (int i = i0; i < i1; ++i) {
    const float q = (float)((double)p.probs[i] / total);
    p.probs[i] = q;
    psum += (double)q;
  }

  if (p.topp <= 0.0 || p.topp >= 1.0) {
    // sample(): CDF walk scaled by the f64 sum of the probabilities (llama2.ts:368-376)
    double ptotal;
    double cum = block_excl_scan_f64(psum, s_d, &ptotal);
    const double r = p.rand01 * ptotal;
    int hit = 0x7fffffff;
    for (int i = i0; i < i1; ++i) {
      cum += (double)p.probs[i];
      if (r < cum) {
        hit = i;
        break;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hit = min(hit, __shfl_xor_sync(0xffffffffu, hit, o));
    if (lane == 0) s_i[warp] = hit;
    __syncthreads();
    if (tid == 0) {
      int best = s_i[0];
      for (int w = 1; w < 32; ++w) best = min(best, s_i[w]);
      p.next[0] = best == 0x7fffffff ? 0 : best;
    }
    return;
  }

  // sample_topp() (llama2.ts:378-394)
  float cutoff = (float)((1.0 - p.topp) / (double)(V - 1));
  for (int attempt = 0; attempt < 2; ++attempt) {
    __syncthreads();
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += (p.probs[i] >= cutoff) ? 1 : 0;
    int k;
    int at = block_excl_scan_i32(cnt, s_i, &k);
    for (int i = i0; i < i1; ++i) {
      const float q = p.probs[i];
      if (q >= cutoff) {
        p.cand_p[at] = q;
        p.cand_i[at] = i;
        ++at;
      }
    }
    __threadfence_block();
    __syncthreads();
    const float* fin_p = p.sort_p;
    const int* fin_i = p.sort_i;
    if (k <= 2048) {
      // rank = position after the reference's stable sort by descending probability
      for (int j0 = 0; j0 < k; j0 += kSampThreads) {
        const int j = j0 + tid;
        const float pj = j < k ? p.cand_p[j] : 0.f;
        const int ij = j < k ? p.cand_i[j] : 0;
        int rank = 0;
        __syncthreads();
        for (int m = tid; m < k; m += kSampThreads) {
          s_tile_p[m] = p.cand_p[m];
          s_tile_i[m] = p.cand_i[m];
        }
        __syncthreads();
        if (j < k) {
          for (int m = 0; m < k; ++m) {
            const float pm = s_tile_p[m];
            rank += (pm > pj || (pm == pj && s_tile_i[m] < ij)) ? 1 : 0;
          }
          p.sort_p[rank] = pj;
          p.sort_i[rank] = ij;
        }
      }
    } else {
      // stable LSD radix sort, 8 passes of 4 bits on key = ~bits(prob) (ascending key = descending
      // probability; equal probabilities keep their index order because every pass is stable)
      float* src_p = p.cand_p; int* src_i = p.cand_i;
      float* dst_p = p.sort_p; int* dst_i = p.sort_i;
      const int cc = (k + kSampThreads - 1) / kSampThreads;
      const int j0 = tid * cc < k ? tid * cc : k, j1 = (tid + 1) * cc < k ? (tid + 1) * cc : k;
      for (int pass = 0; pass < 8; ++pass) {
        const int shift = pass * 4;
#pragma unroll
        for (int d = 0; d < 16; ++d) s_cnt[d * kSampThreads + tid] = 0;
        for (int j = j0; j < j1; ++j) {
          const unsigned key = ~__float_as_uint(src_p[j]);
          s_cnt[((key >> shift) & 15u) * kSampThreads + tid] += 1;
        }
        __syncthreads();
        // exclusive scan of the flattened [digit][thread] counters; thread t owns entries [16t, 16t+16)
        int loc[16], sum16 = 0;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          loc[e] = s_cnt[tid * 16 + e];
          sum16 += loc[e];
        }
        int tot;
        int base = block_excl_scan_i32(sum16, s_i, &tot);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          s_cnt[tid * 16 + e] = base;
          base += loc[e];
        }
        __syncthreads();
        for (int j = j0; j < j1; ++j) {
          const float q = src_p[j];
          const unsigned key = ~__float_as_uint(q);
          const int at2 = s_cnt[((key >> shift) & 15u) * kSampThreads + tid]++;
          dst_p[at2] = q;
          dst_i[at2] = src_i[j];
        }
        __syncthreads();
        float* tp = src_p; src_p = dst_p; dst_p = tp;
        int* ti = src_i; src_i = dst_i; dst_i = ti;
      }
      fin_p = src_p;
      fin_i = src_i;
    }
    __threadfence_block();
    __syncthreads();
    // The two walks of llama2.ts:382-393 over the sorted candidates.  Up to 2048 candidates one
    // thread walks them exactly like the reference; above that (flat distributions) the
    // prefix sums are formed per thread chunk + block scan, like sample() above.
    if (k <= 2048) {
      if (tid == 0) {
        double cum = 0.0;
        int last = 0;
        bool found = false;
        for (int i = 0; i < k; ++i) {
          cum += (double)fin_p[i];
          if (cum > p.topp) {
            last = i;
            found = true;
            break;
          }
        }
        s_retry = (!found && k < V) ? 1 : 0;  // the reference's walk would go on into smaller entries
        if (!s_retry) {
          const double r = p.rand01 * cum;
          double c2 = 0.0;
          int res = 0;
          for (int i = 0; i < last; ++i) {
            c2 += (double)fin_p[i];
            if (r < c2) {
              res = fin_i[i];
              break;
            }
          }
          p.next[0] = res;
        }
      }
      __syncthreads();
    } else {
      const int cc = (k + kSampThreads - 1) / kSampThreads;
      const int j0 = tid * cc < k ? tid * cc : k, j1 = (tid + 1) * cc < k ? (tid + 1) * cc : k;
      double csum = 0.0;
      for (int j = j0; j < j1; ++j) csum += (double)fin_p[j];
      double tot;
      const double pre = block_excl_scan_f64(csum, s_d, &tot);
      // first index whose inclusive prefix exceeds topp
      int hit = 0x7fffffff;
      double cum = pre, cum_at = 0.0;
      for (int j = j0; j < j1; ++j) {
        cum += (double)fin_p[j];
        if (cum > p.topp) {
          hit = j;
          cum_at = cum;
          break;
        }
      }
      int wmin = hit;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wmin = min(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
      if (lane == 0) s_i[warp] = wmin;
      __syncthreads();
      int last = s_i[0];
      for (int w = 1; w < 32; ++w) last = min(last, s_i[w]);
      __syncthreads();
      const bool found = last != 0x7fffffff;
      if (found && hit == last) s_d[0] = cum_at;   // the owner publishes cumProb at lastIdx
      if (!found && tid == 0) s_d[0] = tot;
      __syncthreads();
      const double cum_prob = s_d[0];
      __syncthreads();
      if (!found && k < V) {
        if (tid == 0) s_retry = 1;
      } else {
        if (tid == 0) s_retry = 0;
        const int lim = found ? last : 0;          // exclusive walk i < lastIdx
        const double r = p.rand01 * cum_prob;
        int h2 = 0x7fffffff;
        double c2 = pre;
        for (int j = j0; j < j1 && j < lim; ++j) {
          c2 += (double)fin_p[j];
          if (r < c2) {
            h2 = j;
            break;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) h2 = min(h2, __shfl_xor_sync(0xffffffffu, h2, o));
        if (lane == 0) s_i[warp] = h2;
        __syncthreads();
        if (tid == 0) {
          int best = s_i[0];
          for (int w = 1; w < 32; ++w) best = min(best, s_i[w]);
          p.next[0] = best == 0x7fffffff ? 0 : fin_i[best];
        }
      }
      __syncthreads();
    }
    if (!s_retry) break;
    cutoff = -1.f;  // every entry is a candidate
  }
}

}  // namespace l2b
