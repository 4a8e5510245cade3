// common.cuh -- PTX wrappers and small device helpers shared by every kernel.
// sm_100a only: bulk async copies (UBLKCP), mbarriers, clusters/DSMEM, PDL.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace l2b {

constexpr int kWarp = 32;

// ---- global loads ---------------------------------------------------------
// Weight streams are read exactly once per decode step: bypass L1 and attach a
// runtime L2 policy -- evict_first for models larger than the 126 MB L2 (so the
// stream does not push activations / KV rows out), evict_normal for models that
// fit (they then stay L2-resident from step to step).
__device__ __forceinline__ uint64_t make_l2_policy(bool evict_first) {
  uint64_t pf, pn;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pn));
  return evict_first ? pf : pn;
}
__device__ __forceinline__ float4 ldg_stream(const float4* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}

// Activations written by the previous kernel: loads that are coherent at GPU scope (served
// by L2; never .nc, never a possibly stale L1 line): with programmatic dependent launch the
// producer may still be running when this kernel starts, and with the software hand-over
// below it has not formally completed when the data is read.
__device__ __forceinline__ float ld_act(const float* p) {
  float r;
  asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float4 ld_act4(const float4* p) {
  float4 r;
  asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ int ld_act_i32(const int* p) {
  int r;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
// system-scope acquire/release used by the tensor-parallel exchange flags
__device__ __forceinline__ int ld_acquire_sys_i32(const int* p) {
  int r;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_release_sys_i32(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_f32(float* p, float v) {
  asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ long long gtimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- programmatic dependent launch ------------------------------------------
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void griddep_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- software hand-over between consecutive kernels of a step ------------------
// Measured on B200 (tools/gemv_timeline.py): after the last CTA of a kernel exits, its
// programmatic dependent only returns from griddepcontrol.wait 3.5-4.5 us later (grid
// completion + flush + release), 129 times per Llama-2-7B token.  A kernel of the batch-1
// decode chain therefore never executes griddepcontrol.wait; it was launched early anyway
// (programmatic stream serialization) and waits here until every CTA of its predecessor has
// bumped that predecessor's counter -- the arrive/poll half of a grid barrier (~1.3 us).
// Counters are zeroed by a memset node at the start of the step's graph.
__device__ __forceinline__ void soft_signal(int* ctr) {
  __syncthreads();  // every thread's results are written ...
  if (threadIdx.x == 0) {
    __threadfence();  // ... and visible at GPU scope before the arrival is
    atomicAdd(ctr, 1);
  }
}
__device__ __forceinline__ void soft_wait(const int* ctr, int target) {
  if (threadIdx.x == 0) {
    int v;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// ---- warp reductions --------------------------------------------------------
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- shared-memory addresses, mbarriers, bulk async copy ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared (SASS: UBLKCP), completion on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// asynchronous bulk prefetch of a global range into L2 (no destination, no completion event)
__device__ __forceinline__ void prefetch_l2_bulk(const void* g, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}
// generic-proxy writes -> visible to the async proxy (TMA / tcgen05 reads of smem)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- clusters / distributed shared memory -------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// split form: arrive early, wait later (e.g. "every CTA of the cluster has started" before the first
// store into a peer's shared memory, without stalling at kernel entry)
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (a shared variable of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ float dsmem_ld_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double dsmem_ld_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

}  // namespace l2b
