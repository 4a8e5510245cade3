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
// index asc) -- the order of the reference's stable sort -- and then walked SEQUENTIALLY by one
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

__global__ void __launch_bounds__(kSampThreads) sample_kernel(const __grid_constant__ SampleParams p) {
  __shared__ double s_d[33];
  __shared__ int s_i[33];
  __shared__ float s_f[32];
  __shared__ float s_tile_p[2048];
  __shared__ int s_tile_i[2048];
  __shared__ int s_retry, s_k;
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
    // rank = position after the reference's stable sort by descending probability
    for (int j0 = 0; j0 < k; j0 += kSampThreads) {
      const int j = j0 + tid;
      const float pj = j < k ? p.cand_p[j] : 0.f;
      const int ij = j < k ? p.cand_i[j] : 0;
      int rank = 0;
      for (int m0 = 0; m0 < k; m0 += 2048) {
        __syncthreads();
        for (int m = tid; m < 2048 && m0 + m < k; m += kSampThreads) {
          s_tile_p[m] = p.cand_p[m0 + m];
          s_tile_i[m] = p.cand_i[m0 + m];
        }
        __syncthreads();
        const int lim = (k - m0) < 2048 ? (k - m0) : 2048;
        if (j < k)
          for (int m = 0; m < lim; ++m) {
            const float pm = s_tile_p[m];
            rank += (pm > pj || (pm == pj && s_tile_i[m] < ij)) ? 1 : 0;
          }
      }
      if (j < k) {
        p.sort_p[rank] = pj;
        p.sort_i[rank] = ij;
      }
    }
    __threadfence_block();
    __syncthreads();
    if (tid == 0) {
      double cum = 0.0;
      int last = 0;
      bool found = false;
      for (int i = 0; i < k; ++i) {
        cum += (double)p.sort_p[i];
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
          c2 += (double)p.sort_p[i];
          if (r < c2) {
            res = p.sort_i[i];
            break;
          }
        }
        p.next[0] = res;
      }
      s_k = k;
    }
    __syncthreads();
    if (!s_retry) break;
    cutoff = -1.f;  // every entry is a candidate
  }
}

}  // namespace l2b
