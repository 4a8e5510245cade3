// batch_gemm.cuh -- batched decode: B independent sequences turn every matmul()
// of llama2.ts:196-203 into a dense GEMM  C[M][B] = W[M][K] * X[B][K]^T  that runs
// on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// The reference computes in fp32 storage / f64 accumulation; a single TF32 pass
// puts 10k-20k of 32000 logits outside the 1e-4 tolerance,
// so every product is formed as the 3xTF32 split
//       w*x ~= w_hi*x_hi + w_hi*x_lo + w_lo*x_hi        (fp32 accumulate in TMEM)
// with  v_hi = v rounded to nearest at 10 mantissa bits (exactly representable in TF32) and
// v_lo = v - v_hi (exact in fp32, |v_lo| <= 2^-11 |v|, symmetric in sign).
//
// The tensor core adds into its fp32 accumulator with truncation, a bias of ~half an ulp
// per tcgen05.mma that grows with the length of the accumulation chain (measured: logits
// error proportional to K per split).  Chains are kept short three ways: the two small
// terms go to their own accumulator (its truncation error is 2^-11 smaller), the
// accumulators rotate over G copies in TMEM (512 columns: G = 4 for N = 32, 2 for N = 64, 1
// above), and k is split over work items; all partial accumulators are added with
// round-to-nearest fp32 in the epilogues.
//
// For N <= 128 the two products that share the operand w_hi are issued as ONE MMA of width
// 2N against the stacked activation tile [x_hi ; x_lo] (they are adjacent in shared memory),
// landing in adjacent accumulators (main | small); w_lo*x_hi is a second MMA of width N into
// the small one.  At small N the MMA is bound by its shared-memory operand reads, and this
// reads the 4 KB weight operand twice instead of three times per K=8 step.
// Weights arrive ONCE from HBM as fp32 through TMA bulk copies.  They are kept in a second,
// TILE-MAJOR copy ([tile][k-block][128 rows][32 floats], rows pre-swizzled to the
// SWIZZLE_128B image tcgen05.mma expects) so that every 16 KB operand tile is ONE contiguous
// cp.async.bulk (no tensor map, no per-row requests).  Four
// "splitter" warps derive the hi/lo tiles in shared memory in place (an elementwise
// rewrite, so the swizzled layout is preserved), the activations are pre-split by
// the epilogue kernel that produced them.
//
// Warp roles of the 384-thread CTA (one CTA per SM, persistent over work items):
//   warp 0      TMA producer   (cp.async.bulk.tensor into the deep landing ring)
//   warp 1      MMA issuer     (one elected lane, 12 x tcgen05.mma per 32-float k-block)
//   warp 2      TMEM allocator
//   warps 4-7   splitter       (hi/lo tiles of the weights)
//   warps 8-11  epilogue       (tcgen05.ld -> partial sums in global memory); for N <= 128
//               the accumulators are double-buffered in TMEM so the next item's MMAs run
//               while this one drains
// A work item is (128-row tile of W, k-split); partial sums go to P[split][b][m] and
// are reduced in a fixed order by the fused elementwise kernels at the end of this
// file (RoPE + KV write, residual + rmsnorm, SwiGLU, logits + argmax), which also
// emit the pre-split activations of the next GEMM.
#pragma once
#include <cuda.h>
#include <math.h>

#include "common.cuh"
#include "decode_kernels.cuh"

namespace l2b {

constexpr int kGemmThreads = 384;
constexpr int kGemmThreadsTmemA = 512;  // + a second splitter group (warps 12-15)
constexpr int kBM = 128;                    // weight rows per tile (UMMA M)
constexpr int kBK = 32;                     // floats per k-block: one 128-byte swizzle row
constexpr int kTileA = kBM * kBK * 4;       // 16 KB
constexpr uint32_t kHiMask = 0xFFFFE000u;   // keeps sign, exponent and 10 mantissa bits (TF32)

// round-to-nearest split v = hi + lo with hi exactly representable in TF32
__device__ __forceinline__ uint32_t tf32_hi_bits(uint32_t bits) { return (bits + 0x1000u) & kHiMask; }

struct GemmParams {
  float* P;       // partial sums [S][B][M]
  int M, K;       // weight rows, reduction length
  int S;          // k splits
  int B;          // sequences (columns) written out
  int n0;         // first column of this launch (batches > 256 run in column groups)
  int tiles_m;    // ceil(M / 128)
  int kblocks;    // ceil(K / 32)
  const float* Wt; // weights, tile-major + pre-swizzled: [tiles_m][kblocks][128 rows][32 floats]
  const float* Xh; // pre-split activations, k-block-major + pre-swizzled: [K/32][npad][32 floats]
  const float* Xl;
  int npad;       // rows (sequences, padded) per k-block in Xh / Xl
  int dl;         // landing-ring depth (raw weight tiles in flight)
  int dop;        // operand-ring depth
  long long* dbg; // optional timeline of CTA 0 (clock64 per role per k-block), nullptr = off
  int rewrite_hi; // 1: weights' hi part rounded to nearest and rewritten in shared memory;
                  // 0: hi = hardware truncation of the raw tile (saves 16 KB of st.shared per k-block)
};

// One lane of a CONVERGED warp (cute::elect_one_sync).  Unlike `if (lane == 0)`, the
// compiler knows the operands computed by the whole warp are uniform and feeds tcgen05 /
// bulk-copy instructions from uniform registers directly; with `lane == 0` it wrapped every
// tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA waterfall loop (~70 cycles per MMA,
// measured with the kernel timeline: the MMA issuer was the bottleneck at small N).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], both operands K-major SWIZZLE_128B, TF32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem) {
  const uint32_t a = smem_u32(smem);
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// Ring helper: slot index + phase parity of a circular buffer of `depth` slots.
struct Ring {
  int i, ph, depth;
  __device__ __forceinline__ Ring(int d) : i(0), ph(0), depth(d) {}
  __device__ __forceinline__ void next() {
    if (++i == depth) { i = 0; ph ^= 1; }
  }
};

// N = padded number of sequences handled by one MMA (UMMA N): 32, 64, 128 or 256
//
// Shared memory holds two rings.  The LANDING ring (p.dl slots) receives everything TMA
// brings -- the raw fp32 weight tile (16 KB) and the pre-split (hi, lo) activation tiles of
// the same k-block -- and is deep, so that many tiles are in flight in HBM / L2 at once; the
// splitter turns a landed weight tile into its TF32-exact hi part IN PLACE.  The OPERAND
// ring (p.dop slots of 16 KB) holds the only thing that is produced on the SM: the lo part
// of the weight tile.  Only the short split -> MMA -> commit chain is serialised per
// operand slot; the long HBM latency is covered by the landing ring.
template <int N>
__global__ void __launch_bounds__(kGemmThreads, 1)
l2b_tc3x_matmul_kernel(const __grid_constant__ GemmParams p) {
  constexpr int kTileX = N * kBK * 4;
  // N <= 128: the activation tiles travel with the weight tile in the landing ring (deep
  // prefetch of everything; HBM/L2-latency regime).  N = 256 is tensor-bound and its 64 KB of
  // activation tiles per k-block would leave room for only two landing slots, so there they
  // live in the (2-deep) operand ring and have their own producer thread (warp 3).
  constexpr bool kXInLanding = N <= 128;
  constexpr int kLandSlot = kTileA + (kXInLanding ? 2 * kTileX : 0);   // A (raw -> hi) [| X_hi | X_lo]
  constexpr int kOpSlot = kTileA + (kXInLanding ? 0 : 2 * kTileX);     // A_lo [| X_hi | X_lo]
  constexpr int kMaxDepth = 12;
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  constexpr bool kWide = N <= 128;                     // stacked [x_hi ; x_lo] MMA of width 2N
  constexpr uint32_t kIdesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | (8u << 24);
  constexpr int G = N == 32 ? 4 : (N == 64 ? 2 : 1);   // (main | small) accumulator pairs per set
  constexpr int kSetCols = G * 2 * N;                  // 256 (N <= 128) or 512 (N = 256)
  constexpr int kSets = 512 / kSetCols;                // accumulator sets: 2, or 1 for N = 256
  constexpr uint32_t kTmemCols = 512u;

  extern __shared__ __align__(1024) unsigned char gsm[];
  __shared__ __align__(8) uint64_t land_full[kMaxDepth], land_empty[kMaxDepth];
  __shared__ __align__(8) uint64_t op_empty[4], split_done[4], op_xfull[4];
  __shared__ __align__(8) uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;

  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int DL = p.dl, DO = p.dop;
  // dynamic shared memory is only guaranteed 16-byte aligned: round up to the swizzle atom
  unsigned char* sm = gsm + ((1024u - (smem_u32(gsm) & 1023u)) & 1023u);
  unsigned char* land = sm;                             // DL x kLandSlot
  unsigned char* ops = sm + (size_t)DL * kLandSlot;     // DO x kOpSlot

  if (threadIdx.x == 0) {
    for (int s = 0; s < DL; ++s) {
      mbar_init(&land_full[s], 1);
      mbar_init(&land_empty[s], 1);
    }
    for (int s = 0; s < DO; ++s) {
      mbar_init(&op_empty[s], 1);
      mbar_init(&op_xfull[s], 1);
      mbar_init(&split_done[s], 4);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int n_items = p.tiles_m * p.S;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    // Weights never depend on the previous kernel: the first DL weight tiles are requested
    // BEFORE griddep_wait(), their activation tiles right after it (same mbarrier, the
    // expected byte count covers both).
    if (lane == 0) {
      struct Seq {  // the (tile, k-block) sequence of this CTA
        int it, kb, kb1, mt;
        const GemmParams& p;
        int n_items, stride;
        __device__ Seq(const GemmParams& p_, int first, int n, int st) : p(p_), n_items(n), stride(st) {
          it = first;
          load();
        }
        __device__ void load() {
          if (it < n_items) {
            const int split = it / p.tiles_m;
            mt = it - split * p.tiles_m;
            kb = (int)(((long long)p.kblocks * split) / p.S);
            kb1 = (int)(((long long)p.kblocks * (split + 1)) / p.S);
          }
        }
        __device__ bool done() const { return it >= n_items; }
        __device__ void next() {
          if (++kb >= kb1) {
            it += stride;
            load();
          }
        }
      };
      Seq a(p, blockIdx.x, n_items, gridDim.x), x(p, blockIdx.x, n_items, gridDim.x);
      Ring ra(DL), rx(DL);
      int ahead = 0;
      while (!a.done() && ahead < DL) {  // prologue: weights only
        mbar_arrive_expect_tx(&land_full[ra.i], kLandSlot);
        bulk_g2s(land + (size_t)ra.i * kLandSlot, p.Wt + ((size_t)(a.mt) * p.kblocks + (a.kb)) * (kBM * kBK), kTileA, &land_full[ra.i]);
        ra.next();
        a.next();
        ++ahead;
      }
      if (kXInLanding) {
        griddep_wait();  // X was written by the previous kernel
        for (int j = 0; j < ahead; ++j) {
          unsigned char* slot = land + (size_t)rx.i * kLandSlot;
          bulk_g2s(slot + kTileA, p.Xh + ((size_t)(x.kb) * p.npad + p.n0) * kBK, kTileX, &land_full[rx.i]);
          bulk_g2s(slot + kTileA + kTileX, p.Xl + ((size_t)(x.kb) * p.npad + p.n0) * kBK, kTileX, &land_full[rx.i]);
          rx.next();
          x.next();
        }
      }
      while (!a.done()) {  // steady state
        mbar_wait(&land_empty[ra.i], ra.ph ^ 1);
        unsigned char* slot = land + (size_t)ra.i * kLandSlot;
        mbar_arrive_expect_tx(&land_full[ra.i], kLandSlot);
        bulk_g2s(slot, p.Wt + ((size_t)(a.mt) * p.kblocks + (a.kb)) * (kBM * kBK), kTileA, &land_full[ra.i]);
        if (kXInLanding) {
          bulk_g2s(slot + kTileA, p.Xh + ((size_t)(a.kb) * p.npad + p.n0) * kBK, kTileX, &land_full[ra.i]);
          bulk_g2s(slot + kTileA + kTileX, p.Xl + ((size_t)(a.kb) * p.npad + p.n0) * kBK, kTileX, &land_full[ra.i]);
        }
        ra.next();
        a.next();
      }
    }
  } else if (warp == 3 && !kXInLanding) {
    // ---------------- activation producer (N = 256) ----------------
    if (lane == 0) {
      griddep_wait();  // X was written by the previous kernel
      Ring ro(DO);
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int split = it / p.tiles_m;
        const int kb0 = (int)(((long long)p.kblocks * split) / p.S);
        const int kb1 = (int)(((long long)p.kblocks * (split + 1)) / p.S);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&op_empty[ro.i], ro.ph ^ 1);
          unsigned char* slot = ops + (size_t)ro.i * kOpSlot;
          mbar_arrive_expect_tx(&op_xfull[ro.i], 2 * kTileX);
          bulk_g2s(slot + kTileA, p.Xh + ((size_t)(kb) * p.npad + p.n0) * kBK, kTileX, &op_xfull[ro.i]);
          bulk_g2s(slot + kTileA + kTileX, p.Xl + ((size_t)(kb) * p.npad + p.n0) * kBK, kTileX, &op_xfull[ro.i]);
          ro.next();
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (whole warp, one elected lane issues) ----------------
    {
      Ring rl(DL), ro(DO);
      int li = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
        const int split = it / p.tiles_m;
        const int kb0 = (int)(((long long)p.kblocks * split) / p.S);
        const int kb1 = (int)(((long long)p.kblocks * (split + 1)) / p.S);
        const int set = li % kSets;
        mbar_wait(&tmem_empty_bar[set], ((li / kSets) & 1) ^ 1);  // epilogue drained this set
        tc_fence_after();
        const uint32_t tset = tmem_base + (uint32_t)(set * kSetCols);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int kk = kb - kb0;
          const uint32_t d_main = tset + (uint32_t)((kk % G) * 2 * N);     // rotating pair: main | small
          const uint32_t d_small = d_main + (uint32_t)N;
          const uint32_t acc0 = kk >= G ? 1u : 0u;                         // first use overwrites
          mbar_wait(&split_done[ro.i], ro.ph);   // hi (landing slot) and lo (operand slot) written
          if (!kXInLanding) mbar_wait(&op_xfull[ro.i], ro.ph);
          tc_fence_after();
          unsigned char* ls = land + (size_t)rl.i * kLandSlot;
          unsigned char* os = ops + (size_t)ro.i * kOpSlot;
          unsigned char* xs = kXInLanding ? ls + kTileA : os + kTileA;
          const uint64_t dAh = umma_desc_sw128(ls), dAl = umma_desc_sw128(os);
          const uint64_t dXh = umma_desc_sw128(xs);
          const uint64_t dXl = umma_desc_sw128(xs + kTileX);
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // 32 bytes per K=8 step
            const uint32_t acc = k != 0 ? 1u : acc0;
            if (!elect_one()) {
            } else if (kWide) {
              tc_mma_tf32(d_main, dAh + adv, dXh + adv, kIdesc2, acc);   // w_hi * [x_hi ; x_lo]
              tc_mma_tf32(d_small, dAl + adv, dXh + adv, kIdesc, 1u);    // w_lo * x_hi
            } else {
              tc_mma_tf32(d_small, dAl + adv, dXh + adv, kIdesc, acc);
              tc_mma_tf32(d_small, dAh + adv, dXl + adv, kIdesc, 1u);
              tc_mma_tf32(d_main, dAh + adv, dXh + adv, kIdesc, acc);
            }
            __syncwarp();
          }
          if (elect_one()) {
            tc_commit(&land_empty[rl.i]);  // both slots are free once these MMAs have read them
            tc_commit(&op_empty[ro.i]);
          }
          __syncwarp();
          rl.next();
          ro.next();
        }
        if (elect_one()) tc_commit(&tmem_full_bar[set]);  // accumulators complete
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------- splitter ----------------
    const int t = threadIdx.x - 128;  // 0..127
    const bool rewrite_hi = p.rewrite_hi != 0;
    Ring rl(DL), ro(DO);
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int split = it / p.tiles_m;
      const int kb0 = (int)(((long long)p.kblocks * split) / p.S);
      const int kb1 = (int)(((long long)p.kblocks * (split + 1)) / p.S);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&op_empty[ro.i], ro.ph ^ 1);  // the lo slot is free again
        mbar_wait(&land_full[rl.i], rl.ph);     // the raw tile has landed
        uint4* ah = reinterpret_cast<uint4*>(land + (size_t)rl.i * kLandSlot);
        uint4* al = reinterpret_cast<uint4*>(ops + (size_t)ro.i * kOpSlot);
#pragma unroll
        for (int i = 0; i < kTileA / 16 / 128; ++i) {
          const int c = t + i * 128;
          const uint4 v = ah[c];
          uint4 h, l;
          if (rewrite_hi) {  // round-to-nearest hi, written back in place
            h.x = tf32_hi_bits(v.x); h.y = tf32_hi_bits(v.y); h.z = tf32_hi_bits(v.z); h.w = tf32_hi_bits(v.w);
          } else {           // the tensor core ignores the low 13 bits: hi = truncation, tile untouched
            h.x = v.x & kHiMask; h.y = v.y & kHiMask; h.z = v.z & kHiMask; h.w = v.w & kHiMask;
          }
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
          if (rewrite_hi) ah[c] = h;
          al[c] = l;
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_done[ro.i]);
        rl.next();
        ro.next();
      }
    }
  } else if (warp >= 8) {
    // ---------------- epilogue: TMEM -> registers -> P[split][b][m] ----------------
    const int wq = warp & 3;  // TMEM lane quadrant this warp may read
    int li = 0;
    griddep_wait();  // P may still be read by the previous kernel
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
      const int split = it / p.tiles_m, mt = it - split * p.tiles_m;
      const int kb0 = (int)(((long long)p.kblocks * split) / p.S);
      const int kb1 = (int)(((long long)p.kblocks * (split + 1)) / p.S);
      const int set = li % kSets;
      mbar_wait(&tmem_full_bar[set], (li / kSets) & 1);
      tc_fence_after();
      const int m = mt * kBM + wq * 32 + lane;
      float* out = p.P + ((size_t)split * p.B + p.n0) * p.M + m;
      const int used = (kb1 - kb0) < G ? (kb1 - kb0) : G;  // main accumulators written by this item
      const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(set * kSetCols);
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 32) {
        if (p.n0 + c0 >= p.B) break;
        uint32_t r[32];
        float sum[32], small[32];
        tmem_ld32(lane_base + (uint32_t)c0, r);        // pair 0: main
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[j] = __uint_as_float(r[j]);
        tmem_ld32(lane_base + (uint32_t)(N + c0), r);  // pair 0: small
#pragma unroll
        for (int j = 0; j < 32; ++j) small[j] = __uint_as_float(r[j]);
#pragma unroll 1
        for (int g = 1; g < used; ++g) {
          tmem_ld32(lane_base + (uint32_t)(g * 2 * N + c0), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(r[j]);
          tmem_ld32(lane_base + (uint32_t)(g * 2 * N + N + c0), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) small[j] += __uint_as_float(r[j]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[j] += small[j];
        if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (p.n0 + c0 + j < p.B) out[(size_t)(c0 + j) * p.M] = sum[j];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[set]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------
// Variant for N <= 128 with the WEIGHT operand in tensor memory.
//
// ncu on the kernel above at N = 32 (7B, 32 sequences): DRAM 43 %, tensor pipe 17 % -- it is
// bound by shared-memory bandwidth: per 16 KB weight tile the MMAs re-read 32 KB of weight
// operand (w_hi, w_lo) + 12 KB of activations from shared memory, on top of the TMA write
// and the splitter's read + write.  Here the splitter (one thread per weight row) reads the
// landed tile once and writes w_hi / w_lo with tcgen05.st into a 4-slot ring in TMEM (64
// columns per slot); tcgen05.mma takes A from TMEM, so shared memory only sees the TMA write,
// one read of the tile, and the small activation operand.
// TMEM columns: [0,256) accumulators (2 sets x G pairs of (main | small)), [256,512) A ring.
template <int N>
__global__ void __launch_bounds__(kGemmThreadsTmemA, 1)
l2b_tc3x_tmemA_matmul_kernel(const __grid_constant__ GemmParams p) {
  static_assert(N == 32 || N == 64 || N == 128, "tensor-memory A variant: N <= 128");
  constexpr int kTileX = N * kBK * 4;
  constexpr int kLandSlot = kTileA + 2 * kTileX;   // raw weight tile | X_hi | X_lo
  constexpr int kMaxDepth = 12;
  constexpr int kASlots = 4, kASlotCols = 64, kAFirstCol = 256;
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  constexpr uint32_t kIdesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | (8u << 24);
  constexpr int G = N == 32 ? 2 : 1;               // (main | small) accumulator pairs per set
  constexpr int kSetCols = G * 2 * N;              // 128 (N = 32, 64) or 256 (N = 128)
  constexpr int kSets = N == 128 ? 1 : 2;
  constexpr uint32_t kTmemCols = 512u;

  extern __shared__ __align__(1024) unsigned char gsm[];
  __shared__ __align__(8) uint64_t land_full[kMaxDepth], land_empty[kMaxDepth];
  __shared__ __align__(8) uint64_t a_empty[kASlots], a_full[kASlots];
  __shared__ __align__(8) uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;

  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int DL = p.dl;
  unsigned char* land = gsm + ((1024u - (smem_u32(gsm) & 1023u)) & 1023u);

  if (threadIdx.x == 0) {
    for (int s = 0; s < DL; ++s) {
      mbar_init(&land_full[s], 1);
      mbar_init(&land_empty[s], 1);
    }
    for (int s = 0; s < kASlots; ++s) {
      mbar_init(&a_empty[s], 1);
      mbar_init(&a_full[s], 4);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int n_items = p.tiles_m * p.S;

  if (warp == 0) {
    // ---------------- TMA producer (same as above) ----------------
    if (lane == 0) {
      struct Seq {
        int it, kb, kb1, mt;
        const GemmParams& p;
        int n_items, stride;
        __device__ Seq(const GemmParams& p_, int first, int n, int st) : p(p_), n_items(n), stride(st) {
          it = first;
          load();
        }
        __device__ void load() {
          if (it < n_items) {
            const int split = it / p.tiles_m;
            mt = it - split * p.tiles_m;
            kb = (int)(((long long)p.kblocks * split) / p.S);
            kb1 = (int)(((long long)p.kblocks * (split + 1)) / p.S);
          }
        }
        __device__ bool done() const { return it >= n_items; }
        __device__ void next() {
          if (++kb >= kb1) {
            it += stride;
            load();
          }
        }
      };
      Seq a(p, blockIdx.x, n_items, gridDim.x), x(p, blockIdx.x, n_items, gridDim.x);
      Ring ra(DL), rx(DL);
      int ahead = 0;
      while (!a.done() && ahead < DL) {
        mbar_arrive_expect_tx(&land_full[ra.i], kLandSlot);
        bulk_g2s(land + (size_t)ra.i * kLandSlot, p.Wt + ((size_t)(a.mt) * p.kblocks + (a.kb)) * (kBM * kBK), kTileA, &land_full[ra.i]);
        ra.next();
        a.next();
        ++ahead;
      }
      griddep_wait();
      for (int j = 0; j < ahead; ++j) {
        unsigned char* slot = land + (size_t)rx.i * kLandSlot;
        bulk_g2s(slot + kTileA, p.Xh + ((size_t)(x.kb) * p.npad + p.n0) * kBK, kTileX, &land_full[rx.i]);
        bulk_g2s(slot + kTileA + kTileX, p.Xl + ((size_t)(x.kb) * p.npad + p.n0) * kBK, kTileX, &land_full[rx.i]);
        rx.next();
        x.next();
      }
      int ord = ahead;
      while (!a.done()) {
        mbar_wait(&land_empty[ra.i], ra.ph ^ 1);
        if (p.dbg && blockIdx.x == 0 && ord < 256) p.dbg[ord * 8 + 0] = clock64();
        ++ord;
        unsigned char* slot = land + (size_t)ra.i * kLandSlot;
        mbar_arrive_expect_tx(&land_full[ra.i], kLandSlot);
        bulk_g2s(slot, p.Wt + ((size_t)(a.mt) * p.kblocks + (a.kb)) * (kBM * kBK), kTileA, &land_full[ra.i]);
        bulk_g2s(slot + kTileA, p.Xh + ((size_t)(a.kb) * p.npad + p.n0) * kBK, kTileX, &land_full[ra.i]);
        bulk_g2s(slot + kTileA + kTileX, p.Xl + ((size_t)(a.kb) * p.npad + p.n0) * kBK, kTileX, &land_full[ra.i]);
        ra.next();
        a.next();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: A from tensor memory (whole warp, one elected lane issues) ----
    {
      Ring rl(DL), ro(kASlots);
      int li = 0, mord = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
        const int split = it / p.tiles_m;
        const int kb0 = (int)(((long long)p.kblocks * split) / p.S);
        const int kb1 = (int)(((long long)p.kblocks * (split + 1)) / p.S);
        const int set = li % kSets;
        mbar_wait(&tmem_empty_bar[set], ((li / kSets) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tset = tmem_base + (uint32_t)(set * kSetCols);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int kk = kb - kb0;
          const uint32_t d_main = tset + (uint32_t)((kk % G) * 2 * N);
          const uint32_t d_small = d_main + (uint32_t)N;
          const uint32_t acc0 = kk >= G ? 1u : 0u;
          mbar_wait(&a_full[ro.i], ro.ph);  // w_hi / w_lo of this k-block are in TMEM (=> tile landed)
          if (p.dbg && blockIdx.x == 0 && mord < 256 && lane == 0) p.dbg[mord * 8 + 4] = clock64();
          tc_fence_after();
          unsigned char* ls = land + (size_t)rl.i * kLandSlot;
          const uint64_t dXh = umma_desc_sw128(ls + kTileA);
          const uint32_t a_hi = tmem_base + (uint32_t)(kAFirstCol + ro.i * kASlotCols);
          const uint32_t a_lo = a_hi + 32u;
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
            const uint32_t acc = k != 0 ? 1u : acc0;
            if (elect_one()) {
              asm volatile(
                  "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_main),
                  "r"(a_hi + (uint32_t)(k * 8)), "l"(dXh + adv), "r"(kIdesc2), "r"(acc)
                  : "memory");  // w_hi * [x_hi ; x_lo]
              asm volatile(
                  "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                  "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_small),
                  "r"(a_lo + (uint32_t)(k * 8)), "l"(dXh + adv), "r"(kIdesc), "r"(1u)
                  : "memory");  // w_lo * x_hi
            }
            __syncwarp();
          }
          if (elect_one()) {
            tc_commit(&land_empty[rl.i]);
            tc_commit(&a_empty[ro.i]);
          }
          __syncwarp();
          if (p.dbg && blockIdx.x == 0 && mord < 256 && lane == 0) p.dbg[mord * 8 + 5] = clock64();
          ++mord;
          rl.next();
          ro.next();
        }
        if (elect_one()) tc_commit(&tmem_full_bar[set]);
        __syncwarp();
      }
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 12) {
    // ---------------- splitter: one thread per weight row, shared memory -> TMEM ----------------
    // Two groups of four warps (4-7 and 12-15; warp % 4 selects the TMEM lane quadrant) take
    // alternate k-blocks, so that one group's tcgen05.st latency hides behind the other's loads.
    const int grp = warp >= 12 ? 1 : 0;
    const int row = (warp & 3) * 32 + lane;   // 0..127 == TMEM lane
    int total = 0;                            // k-blocks this CTA processes over all its items
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int split = it / p.tiles_m;
      total += (int)(((long long)p.kblocks * (split + 1)) / p.S) - (int)(((long long)p.kblocks * split) / p.S);
    }
    for (int n = grp; n < total; n += 2) {
      const int ls = n % DL, lph = (n / DL) & 1;
      const int as = n % kASlots, aph = (n / kASlots) & 1;
      mbar_wait(&a_empty[as], aph ^ 1);
      const bool rec = p.dbg && blockIdx.x == 0 && n < 256 && (warp & 3) == 0 && lane == 0;
      if (rec) p.dbg[n * 8 + 1] = clock64();
      tc_fence_after();
      mbar_wait(&land_full[ls], lph);
      if (rec) p.dbg[n * 8 + 2] = clock64();
      const unsigned char* tile = land + (size_t)ls * kLandSlot + (size_t)row * 128;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // logical 16-byte chunk j sits at (j ^ (row & 7)) (SWIZZLE_128B)
        const uint4 v = *reinterpret_cast<const uint4*>(tile + ((j ^ (row & 7)) << 4));
        hi[4 * j + 0] = v.x; hi[4 * j + 1] = v.y; hi[4 * j + 2] = v.z; hi[4 * j + 3] = v.w;
        lo[4 * j + 0] = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & kHiMask));
        lo[4 * j + 1] = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & kHiMask));
        lo[4 * j + 2] = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & kHiMask));
        lo[4 * j + 3] = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & kHiMask));
      }
      const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kAFirstCol + as * kASlotCols);
      tmem_st32(ta, hi);        // the tensor core ignores the low 13 bits: raw words are w_hi
      tmem_st32(ta + 32u, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[as]);
      if (rec) p.dbg[n * 8 + 3] = clock64();
    }
  } else if (warp >= 8 && warp < 12) {
    // ---------------- epilogue (same as above) ----------------
    const int wq = warp & 3;
    int li = 0;
    griddep_wait();
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++li) {
      const int split = it / p.tiles_m, mt = it - split * p.tiles_m;
      const int kb0 = (int)(((long long)p.kblocks * split) / p.S);
      const int kb1 = (int)(((long long)p.kblocks * (split + 1)) / p.S);
      const int set = li % kSets;
      mbar_wait(&tmem_full_bar[set], (li / kSets) & 1);
      tc_fence_after();
      const int m = mt * kBM + wq * 32 + lane;
      float* out = p.P + ((size_t)split * p.B + p.n0) * p.M + m;
      const int used = (kb1 - kb0) < G ? (kb1 - kb0) : G;
      const uint32_t lane_base = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(set * kSetCols);
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 32) {
        if (p.n0 + c0 >= p.B) break;
        uint32_t r[32];
        float sum[32];
        tmem_ld32(lane_base + (uint32_t)c0, r);                       // main accumulators first
#pragma unroll
        for (int j = 0; j < 32; ++j) sum[j] = __uint_as_float(r[j]);
#pragma unroll 1
        for (int g = 1; g < used; ++g) {
          tmem_ld32(lane_base + (uint32_t)(g * 2 * N + c0), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(r[j]);
        }
#pragma unroll 1
        for (int g = 0; g < used; ++g) {                              // then the small terms
          tmem_ld32(lane_base + (uint32_t)(g * 2 * N + N + c0), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(r[j]);
        }
        if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (p.n0 + c0 + j < p.B) out[(size_t)(c0 + j) * p.M] = sum[j];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[set]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols)
                 : "memory");
  }
}

// Row-major [M][K] -> tile-major pre-swizzled copy used by the GEMMs (one-time, at first
// batched use; partial tiles are zero-padded).  One thread per 16-byte chunk.
__global__ void __launch_bounds__(256) l2b_tile_major_kernel(const float* __restrict__ W, float* __restrict__ Wt, int M,
                                                         int K, int tiles_m, int kblocks) {
  const size_t n_chunks = (size_t)tiles_m * kblocks * (kBM * kBK / 4);
  for (size_t q = (size_t)blockIdx.x * 256 + threadIdx.x; q < n_chunks; q += (size_t)gridDim.x * 256) {
    const size_t tile = q / (kBM * kBK / 4);
    const int within = (int)(q - tile * (kBM * kBK / 4));
    const int r = within >> 3, pc = within & 7;
    const int c = pc ^ (r & 7);                      // logical chunk stored at physical position pc
    const int mt = (int)(tile / kblocks), kb = (int)(tile - (size_t)mt * kblocks);
    const int row = mt * kBM + r, col = kb * kBK + c * 4;
    float4 v = f4_zero();
    if (row < M && col < K) v = *reinterpret_cast<const float4*>(W + (size_t)row * K + col);
    reinterpret_cast<float4*>(Wt)[q] = v;
  }
}

// ---------------------------------------------------------------------------
// fused elementwise kernels between the GEMMs (CUDA cores; activations are tiny)
// ---------------------------------------------------------------------------
// Activation operand layout the GEMMs read with ONE contiguous bulk copy per k-block:
// k-block-major [K/32][npad rows][32 floats], each row's eight 16-byte chunks already in
// SWIZZLE_128B order (chunk c of row b at position c ^ (b & 7)) -- exactly the shared-memory
// image tcgen05.mma expects, so no tensor map (and none of its per-row requests) is needed.
__device__ __forceinline__ size_t xsw_index(int b, int j, int npad) {
  return ((size_t)(j >> 5) * npad + b) * 32 + ((((j >> 2) & 7) ^ (b & 7)) << 2) + (j & 3);
}
__device__ __forceinline__ void store_split(float* xh, float* xl, size_t i, float v) {
  const float h = __uint_as_float(tf32_hi_bits(__float_as_uint(v)));
  xh[i] = h;
  xl[i] = v - h;
}

// four consecutive elements (index a multiple of 4: contiguous in the xsw_index layout)
__device__ __forceinline__ void store_split4(float* xh, float* xl, size_t i, float4 v) {
  float4 h, l;
  h.x = __uint_as_float(tf32_hi_bits(__float_as_uint(v.x)));
  h.y = __uint_as_float(tf32_hi_bits(__float_as_uint(v.y)));
  h.z = __uint_as_float(tf32_hi_bits(__float_as_uint(v.z)));
  h.w = __uint_as_float(tf32_hi_bits(__float_as_uint(v.w)));
  l.x = v.x - h.x;
  l.y = v.y - h.y;
  l.z = v.z - h.z;
  l.w = v.w - h.w;
  *reinterpret_cast<float4*>(xh + i) = h;
  *reinterpret_cast<float4*>(xl + i) = l;
}

struct BatVecParams {
  const float* P;        // partial sums [S][B][M] or nullptr
  int S, B, M;
  const float* tok_emb;  // != nullptr: x := embedding row (llama2.ts:211)
  const int* tokp;
  float* x;              // [B][D]
  const float* rms_w;    // rmsnorm weight of the NEXT projection
  float* xh;             // pre-split normalised activations (xsw_index layout)
  float* xl;
  int npad;
  int D;
};

// x += sum_s P[s][b][:]  (accum, llama2.ts:168-170)  or  x := emb[token];
// then rmsnorm (llama2.ts:172-179) -> hi/lo split for the next GEMM.
// 65 of these run per step.  CLUSTER = false: one CTA per sequence (large batches, small rows).
// CLUSTER = true: one cluster of 2/4/8 CTAs per sequence (grid (cluster size, B)), each CTA
// owns a contiguous slice of the row and the sum of squares is completed through distributed
// shared memory in fixed rank order -- with one CTA per sequence a batch of 8-32 kept 8-32 SMs
// busy for ~28 us each time (a quarter of the step).  All loads of a block of 4 float4 per
// thread are issued before the first is used (L2-only loads: the partial sums were written by
// the kernel before this one).
template <bool CLUSTER>
__device__ __forceinline__ void bat_resid_rms_body(const BatVecParams& p) {
  __shared__ double red[8];
  __shared__ double c_part;
  constexpr int KQ = 4;
  const int b = CLUSTER ? blockIdx.y : blockIdx.x, D = p.D, D4 = D >> 2;
  int q0 = 0, q1 = D4;
  uint32_t nrank = 1;
  if (CLUSTER) {
    nrank = cluster_nctarank();
    const int per = (D4 + (int)nrank - 1) / (int)nrank;
    q0 = (int)cluster_ctarank() * per;
    q1 = (q0 + per < D4) ? q0 + per : D4;
  }
  float4* x4 = reinterpret_cast<float4*>(p.x + (size_t)b * D);
  const float4* emb4 =
      p.tok_emb != nullptr ? reinterpret_cast<const float4*>(p.tok_emb + (size_t)ld_act_i32(p.tokp + b) * D) : nullptr;
  const size_t M4 = (size_t)p.M >> 2;
  double ss = 0.0;
  for (int base = q0; base < q1; base += 256 * KQ) {
    float4 xv[KQ], acc[KQ];
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
      const int q = base + (int)threadIdx.x + 256 * k;
      acc[k] = f4_zero();
      xv[k] = f4_zero();
      if (q < q1) xv[k] = emb4 != nullptr ? __ldg(emb4 + q) : __ldcg(x4 + q);
    }
    if (emb4 == nullptr) {
      for (int s = 0; s < p.S; ++s) {
        const float4* P4 = reinterpret_cast<const float4*>(p.P) + ((size_t)s * p.B + b) * M4;
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          const int q = base + (int)threadIdx.x + 256 * k;
          if (q < q1) {
            const float4 t = __ldcg(P4 + q);
            acc[k].x += t.x;
            acc[k].y += t.y;
            acc[k].z += t.z;
            acc[k].w += t.w;
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
      const int q = base + (int)threadIdx.x + 256 * k;
      if (q < q1) {
        float4 v = xv[k];
        if (emb4 == nullptr) {
          v.x = (float)((double)xv[k].x + (double)acc[k].x);
          v.y = (float)((double)xv[k].y + (double)acc[k].y);
          v.z = (float)((double)xv[k].z + (double)acc[k].z);
          v.w = (float)((double)xv[k].w + (double)acc[k].w);
        }
        x4[q] = v;
        ss += (double)v.x * (double)v.x + (double)v.y * (double)v.y + (double)v.z * (double)v.z +
              (double)v.w * (double)v.w;
      }
    }
  }
  ss = warp_sum_f64(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  double tot = 0.0;
  if (CLUSTER) {
    if (threadIdx.x == 0) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w];
      c_part = t;
    }
    cluster_sync_all();
    for (uint32_t r = 0; r < nrank; ++r) tot += dsmem_ld_f64(dsmem_addr(&c_part, r));
  } else {
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
  }
  tot /= (double)D;
  tot = 1.0 / sqrt(1e-5 + tot);
  const float4* w4 = reinterpret_cast<const float4*>(p.rms_w);
  for (int q = q0 + (int)threadIdx.x; q < q1; q += 256) {
    const float4 v = x4[q];  // this thread's own stores
    const float4 w = __ldg(w4 + q);
    float4 o;
    o.x = (float)((double)w.x * (tot * (double)v.x));
    o.y = (float)((double)w.y * (tot * (double)v.y));
    o.z = (float)((double)w.z * (tot * (double)v.z));
    o.w = (float)((double)w.w * (tot * (double)v.w));
    store_split4(p.xh, p.xl, xsw_index(b, 4 * q, p.npad), o);
  }
  if (CLUSTER) cluster_sync_all();  // nobody exits while a peer may still read its partial sum
}
__global__ void __launch_bounds__(256) l2b_bat_resid_rms_kernel(const __grid_constant__ BatVecParams p) {
  griddep_launch_dependents();
  griddep_wait();
  bat_resid_rms_body<false>(p);
}
__global__ void __launch_bounds__(256) l2b_bat_resid_rms_cluster_kernel(const __grid_constant__ BatVecParams p) {
  griddep_launch_dependents();
  griddep_wait();
  bat_resid_rms_body<true>(p);
}

struct BatQkvParams {
  const float* P;  // [S][B][3D]
  int S, B, D, hs, steps;
  const int* posp;
  const float* fcr;
  const float* fci;
  float* q;        // [B][D]
  float* kc;       // layer base [B][H][steps][hs]
  float* vc;
  long long kv_seq_stride;
};

// RoPE (llama2.ts:224-235) + KV-cache write (:238-240); one thread per two row pairs (float4).
__global__ void __launch_bounds__(256) l2b_bat_qkv_epi_kernel(const __grid_constant__ BatQkvParams p) {
  griddep_launch_dependents();
  griddep_wait();
  const int b = blockIdx.y;
  const int quad = blockIdx.x * 256 + threadIdx.x;
  const int M = 3 * p.D;
  const int r = 4 * quad;
  if (r >= M) return;
  // latency chain of this short-lived thread: pos -> freq_cis, overlapped with the partial sums
  const int pos = ld_act_i32(p.posp + b);
  const int seg = r / p.D, i = r - seg * p.D;
  const int h = i / p.hs, c = i - h * p.hs;
  float4 a = f4_zero();
  for (int s0 = 0; s0 < p.S; s0 += 4) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      v[k] = (s0 + k < p.S) ? __ldcg(reinterpret_cast<const float4*>(p.P + ((size_t)(s0 + k) * p.B + b) * M + r))
                            : f4_zero();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (s0 + k < p.S) {
        a.x += v[k].x;
        a.y += v[k].y;
        a.z += v[k].z;
        a.w += v[k].w;
      }
    }
  }
  const size_t row = (size_t)b * p.kv_seq_stride + ((size_t)h * p.steps + pos) * p.hs + c;
  if (seg == 2) {
    *reinterpret_cast<float4*>(p.vc + row) = a;
  } else {
    const float2 fr = __ldg(reinterpret_cast<const float2*>(p.fcr + (size_t)pos * (p.hs / 2) + c / 2));
    const float2 fi = __ldg(reinterpret_cast<const float2*>(p.fci + (size_t)pos * (p.hs / 2) + c / 2));
    float4 o;
    o.x = (float)((double)a.x * (double)fr.x - (double)a.y * (double)fi.x);
    o.y = (float)((double)a.x * (double)fi.x + (double)a.y * (double)fr.x);
    o.z = (float)((double)a.z * (double)fr.y - (double)a.w * (double)fi.y);
    o.w = (float)((double)a.z * (double)fi.y + (double)a.w * (double)fr.y);
    float* dst = seg == 0 ? p.q + (size_t)b * p.D + i : p.kc + row;
    *reinterpret_cast<float4*>(dst) = o;
  }
}

struct BatSwigluParams {
  const float* P;  // [S][B][2F], rows interleaved (2i = w1 row i, 2i+1 = w3 row i)
  int S, B, F;
  float* xh;       // xsw_index layout
  float* xl;
  int npad;
};

// SwiGLU (llama2.ts:284-289) -> pre-split input of the w2 GEMM; four elements per thread
__global__ void __launch_bounds__(256) l2b_bat_swiglu_kernel(const __grid_constant__ BatSwigluParams p) {
  griddep_launch_dependents();
  griddep_wait();
  const int b = blockIdx.y;
  const int i = 4 * (blockIdx.x * 256 + threadIdx.x);
  if (i >= p.F) return;
  float4 u = f4_zero(), w = f4_zero();  // {h1[i], h3[i], h1[i+1], h3[i+1]}, {.. i+2, i+3}
  for (int s0 = 0; s0 < p.S; s0 += 2) {
    float4 v0[2], v1[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float4* src =
          reinterpret_cast<const float4*>(p.P + ((size_t)(s0 + k) * p.B + b) * (2 * (size_t)p.F) + 2 * i);
      v0[k] = (s0 + k < p.S) ? __ldcg(src) : f4_zero();
      v1[k] = (s0 + k < p.S) ? __ldcg(src + 1) : f4_zero();
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (s0 + k < p.S) {
        u.x += v0[k].x; u.y += v0[k].y; u.z += v0[k].z; u.w += v0[k].w;
        w.x += v1[k].x; w.y += v1[k].y; w.z += v1[k].z; w.w += v1[k].w;
      }
    }
  }
  // The exponential in fp32 (expf, <= 2 ulp): this path's inputs already carry the 3xTF32
  // rounding of the GEMM, and 2.8 M fp64 exponentials per launch were what bounded this kernel.
  auto swiglu = [](float h1, float h3) {
    const double hv = (double)h1;
    const float silu = (float)(hv * (1.0 / (1.0 + (double)expf(-h1))));
    return (float)((double)silu * (double)h3);
  };
  const float4 o = make_float4(swiglu(u.x, u.y), swiglu(u.z, u.w), swiglu(w.x, w.y), swiglu(w.z, w.w));
  store_split4(p.xh, p.xl, xsw_index(b, i, p.npad), o);
}

struct BatLogitsParams {
  const float* P;  // [S][B][V]
  int S, B, V;
  float* logits;   // [B][V]
  int* ctl;
  int* next;
  const int* forced;
  int* out_tokens;
};

// logits (llama2.ts:302) + argmax (:364-366) + state advance (:471-504); one CTA per sequence
__global__ void __launch_bounds__(1024) l2b_bat_logits_kernel(const __grid_constant__ BatLogitsParams p) {
  __shared__ float s_v[32];
  __shared__ int s_i[32];
  griddep_launch_dependents();
  griddep_wait();
  const int b = blockIdx.x, V = p.V;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < V; j += 1024) {
    float v = 0.f;
    for (int s = 0; s < p.S; ++s) v += ld_act(p.P + ((size_t)s * p.B + b) * V + j);
    p.logits[(size_t)b * V + j] = v;
    argmax_consider(v, j, bv, bi);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    argmax_consider(ov, oi, bv, bi);
  }
  if ((threadIdx.x & 31) == 0) {
    s_v[threadIdx.x >> 5] = bv;
    s_i[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w) argmax_consider(s_v[w], s_i[w], bv, bi);
    const float l0 = p.logits[(size_t)b * V];
    if (bi == 0x7fffffff || l0 != l0) bi = 0;
    const int B = p.B;
    const int step = ld_act_i32(p.ctl + CTL_STEP);
    int chosen = bi;
    if (ld_act_i32(p.ctl + CTL_USE_FORCED)) {
      const int f = p.forced[(size_t)step * B + b];
      if (f >= 0) chosen = f;
    }
    p.next[b] = bi;
    p.out_tokens[(size_t)step * B + b] = chosen;
    if (ld_act_i32(p.ctl + CTL_ADVANCE)) {
      int* tok = p.ctl + CTL_HDR;
      tok[b] = chosen;
      tok[B + b] = tok[B + b] + 1;
    }
  }
}

// advances the step counter once every sequence has been handled (own launch: the
// per-sequence CTAs above must all have read the old value first)
__global__ void l2b_bat_step_kernel(int* ctl) {
  griddep_launch_dependents();
  griddep_wait();
  if (threadIdx.x == 0 && ctl[CTL_ADVANCE]) ctl[CTL_STEP] = ctl[CTL_STEP] + 1;
}

}  // namespace l2b
