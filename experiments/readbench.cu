// readbench.cu -- how fast can 148 persistent CTAs stream a weight matrix from HBM, as a function of
// the bytes each SM keeps in flight?  (a) the GEMV's scheme: 128-bit loads, register double buffer;
// (b) per-warp rings of shared-memory stages filled by 1-D bulk async copies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/readbench tools/readbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("%s: %s\n", #x, cudaGetErrorString(e));                           \
      exit(1);                                                                 \
    }                                                                          \
  } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// (a) U float4 per thread per tile, two tiles (cur/nxt)
template <int U>
__global__ void __launch_bounds__(512, 1) ldg_kernel(const float4* __restrict__ src, size_t n4, float* out) {
  const size_t per_cta = n4 / gridDim.x;
  const float4* base = src + per_cta * blockIdx.x;
  const size_t tiles = per_cta / (512 * U);
  float4 cur[U], nxt[U];
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < U; ++u) cur[u] = ldg_stream(base + threadIdx.x + 512 * u);
  for (size_t t = 0; t < tiles; ++t) {
    if (t + 1 < tiles) {
#pragma unroll
      for (int u = 0; u < U; ++u) nxt[u] = ldg_stream(base + (t + 1) * 512 * U + threadIdx.x + 512 * u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += cur[u].x + cur[u].y + cur[u].z + cur[u].w;
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = nxt[u];
  }
  if (acc == 12345.678f) out[0] = acc;
}

// (b) per-warp ring: S stages of STG bytes
template <int S, int STG>
__global__ void __launch_bounds__(512, 1) ring_kernel(const float4* __restrict__ src, size_t n4, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[16][S];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t per_warp = n4 / ((size_t)gridDim.x * 16);
  const float4* base = src + per_warp * ((size_t)blockIdx.x * 16 + warp);
  constexpr int STG4 = STG / 16;
  const size_t stages = per_warp / STG4;
  float4* ring = reinterpret_cast<float4*>(smem) + (size_t)warp * S * STG4;
  if (lane == 0) {
    for (int s = 0; s < S; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[warp][s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](size_t st) {
    const int s = (int)(st % S);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[warp][s])), "r"(STG) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(ring + (size_t)s * STG4)),
                 "l"(base + st * STG4), "r"(STG), "r"(smem_u32(&bars[warp][s]))
                 : "memory");
  };
  if (lane == 0)
    for (int s = 0; s < S && (size_t)s < stages; ++s) issue(s);
  float acc = 0.f;
  for (size_t st = 0; st < stages; ++st) {
    const int s = (int)(st % S);
    const uint32_t parity = (uint32_t)((st / S) & 1);
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok)
                   : "r"(smem_u32(&bars[warp][s])), "r"(parity)
                   : "memory");
    const float4* p = ring + (size_t)s * STG4;
#pragma unroll 4
    for (int i = lane; i < STG4; i += 32) {
      const float4 v = p[i];
      acc += v.x + v.y + v.z + v.w;
    }
    __syncwarp();
    if (lane == 0 && st + S < stages) issue(st + S);
  }
  if (acc == 12345.678f) out[0] = acc;
}

// (c) the same two schemes doing the GEMV's arithmetic: rows of N floats, row pairs per warp, x as
// doubles in shared memory, fp64 FMA chains, one warp reduction per pair.
constexpr int N = 4096, N4 = N / 4;
__device__ __forceinline__ double wsum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void fma8(const float4& a, const float4& b, const double2* X, int idx, double (&acc)[2][2], int u) {
  const double2 lo = X[idx], hi = X[N4 + idx];
  double c0 = acc[0][u & 1], c1 = acc[1][u & 1];
  c0 = fma((double)a.x, lo.x, c0); c1 = fma((double)b.x, lo.x, c1);
  c0 = fma((double)a.y, lo.y, c0); c1 = fma((double)b.y, lo.y, c1);
  c0 = fma((double)a.z, hi.x, c0); c1 = fma((double)b.z, hi.x, c1);
  c0 = fma((double)a.w, hi.y, c0); c1 = fma((double)b.w, hi.y, c1);
  acc[0][u & 1] = c0; acc[1][u & 1] = c1;
}
__device__ long long g_cta_ns[3][256];
__device__ int g_cta_sm[3][256];
__device__ int g_rep = 0;
__global__ void __launch_bounds__(512, 1) gemv_ldg_kernel(const float4* __restrict__ W, int rows, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  long long t_start;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
  double2* X = reinterpret_cast<double2*>(smem);
  for (int i = threadIdx.x; i < 2 * N4; i += 512) X[i] = make_double2(1.0, 0.5);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npairs = rows / 2;
  const int p0 = (int)((long long)npairs * blockIdx.x / gridDim.x), p1 = (int)((long long)npairs * (blockIdx.x + 1) / gridDim.x);
  constexpr int TPP = N4 / 128;
  float4 ca[4], cb[4], na[4], nb[4];
  int pair = p0 + warp, jt = 0;
  if (pair < p1) {
    const float4* w0 = W + (size_t)(2 * pair) * N4;
#pragma unroll
    for (int u = 0; u < 4; ++u) { ca[u] = ldg_stream(w0 + lane + 32 * u); cb[u] = ldg_stream(w0 + N4 + lane + 32 * u); }
  }
  double acc[2][2] = {{0, 0}, {0, 0}};
  bool have = pair < p1;
  while (have) {
    int np = pair, nj = jt + 1;
    if (nj == TPP) { nj = 0; np = pair + 16; }
    const bool more = np < p1;
    if (more) {
      const float4* w0 = W + (size_t)(2 * np) * N4 + nj * 128;
#pragma unroll
      for (int u = 0; u < 4; ++u) { na[u] = ldg_stream(w0 + lane + 32 * u); nb[u] = ldg_stream(w0 + N4 + lane + 32 * u); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) fma8(ca[u], cb[u], X, jt * 128 + lane + 32 * u, acc, u);
    if (jt == TPP - 1) {
      const double d0 = wsum(acc[0][0] + acc[0][1]), d1 = wsum(acc[1][0] + acc[1][1]);
      acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0;
      if (lane == 0) { out[2 * pair] = (float)d0; out[2 * pair + 1] = (float)d1; }
    }
    have = more; pair = np; jt = nj;
#pragma unroll
    for (int u = 0; u < 4; ++u) { ca[u] = na[u]; cb[u] = nb[u]; }
  }
  __syncthreads();
  if (threadIdx.x == 0 && g_rep < 3) {
    long long t_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    unsigned sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    g_cta_ns[g_rep][blockIdx.x] = t_end - t_start;
    g_cta_sm[g_rep][blockIdx.x] = (int)sm;
  }
}
__global__ void next_rep() { g_rep++; }
// ring: stage = the two rows' SEG-float segments (two bulk copies), S stages per warp
template <int S, int SEG>
__global__ void __launch_bounds__(512, 1) gemv_ring_kernel(const float4* __restrict__ W, int rows, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[16][S];
  double2* X = reinterpret_cast<double2*>(smem);
  for (int i = threadIdx.x; i < 2 * N4; i += 512) X[i] = make_double2(1.0, 0.5);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int SEG4 = SEG / 4, STG4 = 2 * SEG4, NSEG = N / SEG;
  float4* ring = reinterpret_cast<float4*>(smem + (size_t)N * 8) + (size_t)warp * S * STG4;
  if (lane == 0) {
    for (int s = 0; s < S; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[warp][s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int npairs = rows / 2;
  const int p0 = (int)((long long)npairs * blockIdx.x / gridDim.x), p1 = (int)((long long)npairs * (blockIdx.x + 1) / gridDim.x);
  const int cnt = p1 - p0;
  const int mine = (cnt - warp + 15) / 16;           // pairs p0 + warp, + 16, ...
  const int total = mine > 0 ? mine * NSEG : 0;
  auto issue = [&](int st) {
    const int s = st % S;
    const int pr = p0 + warp + 16 * (st / NSEG), sg = st % NSEG;
    const float4* r0 = W + (size_t)(2 * pr) * N4 + sg * SEG4;
    const uint32_t bar = smem_u32(&bars[warp][s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * SEG * 4) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(ring + (size_t)s * STG4)), "l"(r0), "r"(SEG * 4), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(ring + (size_t)s * STG4 + SEG4)), "l"(r0 + N4), "r"(SEG * 4), "r"(bar) : "memory");
  };
  if (lane == 0)
    for (int s = 0; s < S && s < total; ++s) issue(s);
  double acc[2][2] = {{0, 0}, {0, 0}};
  for (int st = 0; st < total; ++st) {
    const int s = st % S;
    const uint32_t parity = (uint32_t)((st / S) & 1);
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bars[warp][s])), "r"(parity) : "memory");
    const float4* p = ring + (size_t)s * STG4;
    const int sg = st % NSEG;
#pragma unroll
    for (int u = 0; u < SEG4 / 32; ++u) fma8(p[lane + 32 * u], p[SEG4 + lane + 32 * u], X, sg * SEG4 + lane + 32 * u, acc, u);
    __syncwarp();
    if (lane == 0 && st + S < total) issue(st + S);
    if (sg == NSEG - 1) {
      const double d0 = wsum(acc[0][0] + acc[0][1]), d1 = wsum(acc[1][0] + acc[1][1]);
      acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0;
      const int pr = p0 + warp + 16 * (st / NSEG);
      if (lane == 0) { out[2 * pr] = (float)d0; out[2 * pr + 1] = (float)d1; }
    }
  }
}

template <typename F>
float time_it(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t bytes = (size_t)148 * 16 * 1024 * 1024;  // 2.3 GB >> L2, divisible by everything used here
  float4* src;
  float* out;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMalloc(&out, 4));
  CK(cudaMemset(src, 0, bytes));
  const size_t n4 = bytes / 16;
  auto report = [&](const char* name, int inflight_kb, float ms) {
    printf("%-44s %4d KB in flight per SM: %7.1f GB/s\n", name, inflight_kb, bytes / (ms * 1e-3) / 1e9);
    fflush(stdout);
  };
  report("ldg U=4 (GEMV today: 2 rows x 4 float4)", 64, time_it([&] { ldg_kernel<8><<<sms, 512>>>(src, n4, out); }, 5));
  report("ldg U=4", 32, time_it([&] { ldg_kernel<4><<<sms, 512>>>(src, n4, out); }, 5));
  report("ldg U=12", 96, time_it([&] { ldg_kernel<12><<<sms, 512>>>(src, n4, out); }, 5));
#define RING(S, STG)                                                                                          \
  {                                                                                                           \
    const int smem = 16 * S * STG;                                                                            \
    CK(cudaFuncSetAttribute(ring_kernel<S, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));         \
    report("bulk ring S=" #S " stage=" #STG, smem / 1024,                                                      \
           time_it([&] { ring_kernel<S, STG><<<sms, 512, smem>>>(src, n4, out); }, 5));                        \
  }
  RING(2, 2048)
  RING(2, 4096)
  RING(3, 4096)
  RING(4, 2048)
  RING(6, 2048)
  RING(2, 6144)
  RING(3, 2048)
  RING(8, 1024)
  {
    const int rows = (int)(bytes / (N * 4));   // 148 * 1024 rows of 16 KB
    float* o2;
    CK(cudaMalloc(&o2, (size_t)rows * 4));
    auto rep2 = [&](const char* name, float ms) {
      printf("%-60s %7.1f GB/s\n", name, bytes / (ms * 1e-3) / 1e9);
      fflush(stdout);
    };
    CK(cudaFuncSetAttribute(gemv_ldg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, N * 8));
    rep2("GEMV arithmetic, ldg register double buffer (today)", time_it([&] { gemv_ldg_kernel<<<sms, 512, N * 8>>>(src, rows, o2); }, 5));
    {
      // per-CTA duration by SM id over three launches: is the speed of an SM systematic?
      int zero = 0;
      CK(cudaMemcpyToSymbol(g_rep, &zero, sizeof(int)));
      for (int r = 0; r < 3; ++r) {
        gemv_ldg_kernel<<<sms, 512, N * 8>>>(src, rows, o2);
        next_rep<<<1, 1>>>();
      }
      CK(cudaDeviceSynchronize());
      static long long ns[3][256];
      static int sm[3][256];
      CK(cudaMemcpyFromSymbol(ns, g_cta_ns, sizeof(ns)));
      CK(cudaMemcpyFromSymbol(sm, g_cta_sm, sizeof(sm)));
      double by_sm[3][256] = {};
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < sms; ++c) by_sm[r][sm[r][c]] = (double)ns[r][c];
      double mn = 1e30, mx = 0, mean = 0;
      for (int i = 0; i < sms; ++i) { mean += by_sm[0][i]; if (by_sm[0][i] < mn) mn = by_sm[0][i]; if (by_sm[0][i] > mx) mx = by_sm[0][i]; }
      mean /= sms;
      printf("per-CTA duration of one launch: min %.1f us, mean %.1f us, max %.1f us\n", mn / 1e3, mean / 1e3, mx / 1e3);
      // correlation of per-SM durations between launches
      auto corr = [&](int a, int b) {
        double ma = 0, mb = 0;
        for (int i = 0; i < sms; ++i) { ma += by_sm[a][i]; mb += by_sm[b][i]; }
        ma /= sms; mb /= sms;
        double sab = 0, saa = 0, sbb = 0;
        for (int i = 0; i < sms; ++i) { sab += (by_sm[a][i] - ma) * (by_sm[b][i] - mb); saa += (by_sm[a][i] - ma) * (by_sm[a][i] - ma); sbb += (by_sm[b][i] - mb) * (by_sm[b][i] - mb); }
        return sab / sqrt(saa * sbb);
      };
      printf("correlation of per-SM durations: launch 0/1 %.3f, 1/2 %.3f\n", corr(0, 1), corr(1, 2));
      printf("per-SM duration (us), launch 0, by SM id:\n");
      for (int i = 0; i < sms; ++i) printf("%5.0f%s", by_sm[0][i] / 1e3, (i % 16 == 15) ? "\n" : " ");
      printf("\n");
      fflush(stdout);
    }
#define GRING(S, SEG)                                                                                             \
  {                                                                                                               \
    const int smem = N * 8 + 16 * S * 2 * SEG * 4;                                                                \
    CK(cudaFuncSetAttribute(gemv_ring_kernel<S, SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));        \
    rep2("GEMV arithmetic, bulk ring S=" #S " seg=" #SEG " floats/row", time_it([&] { gemv_ring_kernel<S, SEG><<<sms, 512, smem>>>(src, rows, o2); }, 5)); \
  }
    GRING(2, 512)
    GRING(3, 512)
    GRING(2, 1024)
    GRING(4, 256)
  }
  return 0;
}
