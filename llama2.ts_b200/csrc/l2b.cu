// l2b.cu -- host side of libllama2_b200.so: the C ABI declared in
// include/llama2_b200.h.  Owns every device allocation (weights, RunState, KV
// cache), the launch sequence of one decode step (llama2.ts:205-303), its CUDA
// graph, and the device-resident greedy loop (llama2.ts:465-508).
//
// There is NO CPU implementation in this library: without a CUDA device
// l2b_create fails with L2B_ECUDA.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <chrono>

#include <map>
#include <utility>
#include <set>
#include <string>
#include <vector>

#include "../../include/llama2_b200.h"
#include "batch_gemm.cuh"
#include "decode_kernels.cuh"
#ifdef L2B_EXPERIMENTS  // persistent-kernel experiments (measured slower; built only on request)
#include "../../experiments/mega_kernel.cuh"
#include "../../experiments/stream_kernel.cuh"
#endif
#include "fused_qkv_attn.cuh"
#include "sampler.cuh"

#define L2B_API extern "C" __attribute__((visibility("default")))

using namespace l2b;

namespace {

thread_local std::string g_create_error;

struct Options {
  int graph = 1;
  int pdl = 1;
  int f64 = 1;
  int threads = 512;
  int ctas_per_sm = 1;
  int attn_cluster = 0;  // 0 = auto
  int evict_first = -1;  // -1 = auto (weights > L2)
  int tc_min_batch = 3;  // batches >= this run the tcgen05 GEMM path (0 = never).  Measured on 7B: 2 sequences
                         // 5.34 ms/step on the fp64 GEMV path (one weight pass for both); 4 sequences 7.8 ms there
                         // against 5.8 ms on the tensor-core path
  int tc_splits = 0;     // 0 = auto k-split per GEMM
  int tc_rewrite_hi = 0; // see GemmParams::rewrite_hi
  int attn_warp = 2048;  // batched path: one-warp-per-(sequence, head) attention kernel when there are at
                         // least this many (sequence, head) pairs (0 = never); below, the cluster kernel
  int tc_tmem_a = 1;     // N <= 128: weight operand in tensor memory (l2b_tc3x_tmemA_matmul_kernel)
  int l2_prefetch = 262144;  // bytes per CTA prefetched into L2 before griddep_wait (0 = off)
  int attn_prefetch = 0;     // attention kernel prefetches the wo weights into L2 (measured net-negative: it
                             // delays the K/V copies in the same queue; kept as an option)
  int mega = 0;          // batch-1 persistent kernels: 1 = cooperative kernel with grid barriers
                         // (mega_kernel.cuh), 2 = barrier-free streaming kernel (stream_kernel.cuh)
  int fuse_qkv_attn = 1; // batch-1: q/k/v rows and the attention of a layer in one cluster kernel
  int soft_sync = 0;     // batch-1: consecutive kernels hand over through arrival counters instead of
                         // griddepcontrol.wait.  Measured slower (7B 216.5 vs 220.0 tok/s, stories15M 138 vs
                         // 116 us/token): the gap after a kernel is its slowest CTA, not the release latency
  int fuse_cluster = 0;  // CTAs per head of the fused kernel (0 = largest of 8/4/2/1 that fits the SMs)
  int fuse_prefetch = 0; // ... optionally pulling this percentage of wo into L2 while its attention part runs
                         // (measured net-negative: wo 20.0 -> 16.6 us but the fused kernel 46.0 -> 50.6 us)
  int tp_timeout_ms = 20000; // tensor-parallel exchange: bounded spin (see tp_spin_expired)
  int stream_stages = 0; // mega=2: ring stages per warp (0 = as many as shared memory holds)
  int stream_chunks = 0; // mega=2: attention time chunks per head (0 = SMs / heads, at most 8)
                         // (experiment, opt-in: measured slower than graph+PDL so far)
};

}  // namespace

struct l2b_ctx {
  // config (llama2.ts:69-93)
  int D = 0, F = 0, L = 0, H = 0, hs = 0, V = 0, S = 0;
  bool shared_cls = false;
  int device = 0, num_sms = 148;
  int Bmax = 1, steps = 0;
  // tensor parallel (row-sharded projections, one process per GPU)
  int tp_rank = 0, tp_size = 1;
  int Dl = 0, Fl = 0, Vl = 0, Hl = 0;   // this rank's slice of dim / hidden / vocab / heads
  unsigned char* xchg = nullptr;         // exchange block: x | xb | hb | logits | am_val | am_idx | flags
  size_t xchg_bytes = 0;
  size_t off_x = 0, off_xb = 0, off_hb = 0, off_logits = 0, off_amv = 0, off_ami = 0, off_flags = 0;
  unsigned char* peer[kMaxTp] = {nullptr};  // peers' exchange blocks (peer[tp_rank] == xchg)
  bool tp_connected = false;
  int *tp_epoch = nullptr, *tp_ticket = nullptr, *tp_err = nullptr;
  int* h_err = nullptr;
  // weights
  float *tok_emb = nullptr, *rms_att = nullptr, *wqkv = nullptr, *wo = nullptr, *rms_ffn = nullptr,
        *w13 = nullptr, *w2 = nullptr, *rms_final = nullptr, *fcr = nullptr, *fci = nullptr,
        *wcls = nullptr;
  std::vector<unsigned char> uploaded;  // [L2B_T_COUNT][L]
  size_t weight_bytes = 0;
  // run state (llama2.ts:131-163), all on the device
  float *x = nullptr, *xb = nullptr, *q = nullptr, *hb = nullptr, *logits = nullptr;
  float *kc = nullptr, *vc = nullptr;
  int *d_ctl = nullptr, *d_dev = nullptr;  // host-written header / device-only {ticket,next[B]}
  float* blk_val = nullptr;
  int* blk_idx = nullptr;
  int *d_forced = nullptr, *d_out = nullptr;
  int* d_sync = nullptr;      // software hand-over counters, one per kernel of a batch-1 step
  int sync_idx = -1, sync_prev_grid = 0; // chain state while a batch-1 step is being enqueued (-1: off)
  int work_cap = 0;           // entries of d_sync
  unsigned* d_bar = nullptr;  // grid barrier words of the persistent kernel
  uint2* d_ll = nullptr;      // streaming kernel: LL exchange words {value, sequence}; last word: error flag
  size_t ll_words = 0;
  unsigned ll_seq = 1;        // next unused sequence number
  bool ll_used = false;       // a streaming-kernel launch is in flight (finish() reads its error flag)
  float* samp_f = nullptr;    // device sampler scratch: probs | cand_p | sort_p  (3 x vocab)
  int* samp_i = nullptr;      //                         cand_i | sort_i          (2 x vocab)
  // prompt prefill scratch (l2b_prefill): activations of `pf_cap` positions of one sequence
  float *pf_x = nullptr, *pf_xb = nullptr, *pf_q = nullptr;
  int *d_pfctl = nullptr, *h_pfctl = nullptr;
  int pf_cap = 0;
  // batched tensor-core path: pre-split activations [256*groups][D or F], partial sums
  float *XhD = nullptr, *XlD = nullptr, *XhF = nullptr, *XlF = nullptr, *P = nullptr;
  long long* d_dbg2 = nullptr; int gemv_dbg_arm = 0, gemv_dbg_slot = 0;  // GEMV launch timeline (option "gemv_timeline")
  long long* d_dbg = nullptr; // kernel timeline buffer (debug option "gemm_timeline")
  int dbg_arm = 0;
  float* Wt = nullptr;       // tile-major copy of every projection (built at first batched use)
  bool wt_dirty = true;      // weights changed since the copy was built
  bool last_tc = false;      // the last step ran on the tensor-core path (hb is never materialised there)
  std::vector<size_t> wt_off; // float offsets: per layer {qkv, wo, w13, w2}, then cls
  size_t P_floats = 0;
  int Bpad = 0, Smax = 8;
  // host staging (pinned)
  int* h_ctl = nullptr;
  float* h_logits = nullptr;
  int* h_out = nullptr;
  std::vector<int> n_run;  // positions already run per sequence
  // execution
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::map<int, cudaGraphExec_t> graphs;      // key: B
  std::map<int, int> graph_launches;          // kernel nodes per step, key: B
  std::set<const void*> smem_set;
  Options opt;
  float last_ms = 0.f;
  int64_t last_launches = 0;
  int64_t launch_counter = 0;
  // per-class timing (l2b_profile_step)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<int> prof_class;
  // single-process multi-GPU group (l2b_create_multi): this ctx is only a handle, the members own
  // the devices.  tp group: every member is one rank, batch 1.  Otherwise sequence b of the global
  // batch lives on member b / per_kid at local index b % per_kid (weights replicated).
  std::vector<l2b_ctx*> kids;
  int per_kid = 0;
  bool peer_ipc = false;   // peer[] entries were opened with cudaIpcOpenMemHandle (one process per GPU)
  std::string err;
};

namespace {

// Every entry point that selects a device restores the caller's current device on return: the
// host (torch in bench.py, the TS runtime's other native modules) keeps ITS device, also when a
// context group walked over several.
struct DevGuard {
  int prev = -1;
  DevGuard() {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      prev = -1;
      cudaGetLastError();
    }
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int fail(l2b_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  g_create_error = buf;
  return code;
}

#define CU(c, expr)                                                                        \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail((c), L2B_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                     \
  } while (0)

template <typename T>
int dev_alloc(l2b_ctx* c, T** p, size_t count, bool zero) {
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) {
    *p = nullptr;
    cudaGetLastError();
    return fail(c, e == cudaErrorMemoryAllocation ? L2B_ENOMEM : L2B_ECUDA,
                "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
  }
  if (zero) {
    e = cudaMemset(*p, 0, count * sizeof(T));
    if (e != cudaSuccess) return fail(c, L2B_ECUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
  }
  return 0;
}

// tensor-parallel exchange time-out -> device word after the error word, in units of 2^20 clocks
int write_tp_timeout(l2b_ctx* c) {
  if (c->tp_size <= 1 || !c->tp_err) return 0;
  int khz = 1965000;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
  long long units = ((long long)c->opt.tp_timeout_ms * khz) >> 20;
  if (units < 1) units = 1;
  if (units > 0x7fffffffLL) units = 0x7fffffffLL;
  const int u = (int)units;
  CU(c, cudaMemcpy(c->tp_err + 1, &u, sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

void drop_graphs(l2b_ctx* c) {
  for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
  c->graphs.clear();
  c->graph_launches.clear();
}

// ---- kernel selection --------------------------------------------------------
typedef void (*gemv_fn)(const GemvParams);

template <int PRO, int EPI>
gemv_fn pick_gemv(int nb, int threads, bool f64) {
#define L2B_PICK(NB_, T_)                                             \
  if (nb == NB_ && threads == T_)                                     \
    return f64 ? (gemv_fn)l2b_rowpair_matvec_kernel<PRO, EPI, NB_, T_, true>  \
               : (gemv_fn)l2b_rowpair_matvec_kernel<PRO, EPI, NB_, T_, false>;
  L2B_PICK(1, 256)
  L2B_PICK(1, 512)
  L2B_PICK(2, 256)
  L2B_PICK(2, 512)
  L2B_PICK(4, 256)
  L2B_PICK(8, 256)
#undef L2B_PICK
  return nullptr;
}

gemv_fn pick_kernel(int kc, int nb, int threads, bool f64) {
  switch (kc) {
    case L2B_K_QKV: return pick_gemv<PRO_RMS, EPI_QKV>(nb, threads, f64);
    case L2B_K_WO:
    case L2B_K_W2: return pick_gemv<PRO_COPY, EPI_RESID>(nb, threads, f64);
    case L2B_K_W13: return pick_gemv<PRO_RMS, EPI_SWIGLU>(nb, threads, f64);
    case L2B_K_CLS: return pick_gemv<PRO_RMS, EPI_LOGITS>(nb, threads, f64);
  }
  return nullptr;
}

gemv_fn pick_kernel_tp(int kc) {
  switch (kc) {
    case L2B_K_QKV: return (gemv_fn)l2b_rowpair_matvec_kernel<PRO_RMS, EPI_QKV, 1, 512, true, true>;
    case L2B_K_WO:
    case L2B_K_W2: return (gemv_fn)l2b_rowpair_matvec_kernel<PRO_COPY, EPI_RESID, 1, 512, true, true>;
    case L2B_K_W13: return (gemv_fn)l2b_rowpair_matvec_kernel<PRO_RMS, EPI_SWIGLU, 1, 512, true, true>;
    case L2B_K_CLS: return (gemv_fn)l2b_rowpair_matvec_kernel<PRO_RMS, EPI_LOGITS, 1, 512, true, true>;
  }
  return nullptr;
}

int launch(l2b_ctx* c, int kclass, const void* fn, dim3 grid, dim3 block, size_t smem,
           int cluster_x, void** args, cudaStream_t st) {
  if (!c->smem_set.count(fn)) {
    cudaFuncAttributes fa;
    CU(c, cudaFuncGetAttributes(&fa, fn));
    const int max_dyn = 227 * 1024 - (int)fa.sharedSizeBytes;  // static smem counts too
    CU(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn));
    c->smem_set.insert(fn);
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  memset(attrs, 0, sizeof attrs);
  int na = 0;
  if (cluster_x > 1) {
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = cluster_x;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  if (c->opt.pdl && !c->profiling) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  if (c->profiling) {
    cudaEvent_t e;
    CU(c, cudaEventCreate(&e));
    CU(c, cudaEventRecord(e, st));
    c->prof_events.push_back(e);
    c->prof_class.push_back(kclass);
  }
  CU(c, cudaLaunchKernelExC(&cfg, fn, args));
  c->launch_counter++;
  return 0;
}

int pick_nb(const l2b_ctx* c, int B, int n) {
  int nb = 1;
  while (nb < B && nb < kMaxNB) nb *= 2;
  const size_t per = c->opt.f64 ? (size_t)n * 8 : (size_t)n * 4;
  while (nb > 1 && nb * per > 200 * 1024) nb /= 2;
  return nb;
}

// K-chunks of a row pair become separate work units when a CTA owns too few pairs to keep its warps
// streaming (GemvParams::split_k); their lane sums travel through shared memory behind the activation
// vectors: [local pair][chunk][nb][32 lanes] double2 + per-pair counters.  Returns the scratch bytes.
size_t gemv_scratch_bytes(GemvParams& p, int grid, int nb, int threads, size_t vec_bytes) {
  const int n4 = p.n / 4, tpp = (n4 + 32 * kU - 1) / (32 * kU), nch = (tpp + kChunkTiles - 1) / kChunkTiles;
  const int warps = threads / 32;
  p.max_local_pairs = (p.rows / 2 + grid - 1) / grid + 1;
  p.split_k = 0;
  if (nch <= 1) return 0;
  // critical path of a CTA in tile times: whole pairs vs K-chunks as units; split only for a clear win
  // (measured on one GPU, 7B: w2 with 13.8 pairs per 16-warp CTA is faster unsplit, 36.4 vs 38.2 us)
  const int lp = (p.rows / 2 + grid - 1) / grid;
  const int whole = ((lp + warps - 1) / warps) * tpp;
  const int chunks = ((lp * nch + warps - 1) / warps) * kChunkTiles;
  if (chunks * 10 > whole * 8) return 0;
  const size_t bytes = (size_t)p.max_local_pairs * nch * nb * 32 * sizeof(double2) +
                       (((size_t)p.max_local_pairs * sizeof(int)) + 15) / 16 * 16;
  if (vec_bytes + bytes > 200 * 1024) return 0;
  p.split_k = 1;
  return bytes;
}

int launch_gemv(l2b_ctx* c, int kclass, GemvParams& p, int B, cudaStream_t st) {
  if (c->tp_size > 1) {
    gemv_fn fn = pick_kernel_tp(kclass);
    p.B = 1; p.b0 = 0; p.nact = 1; p.dbg = nullptr;
    if (c->gemv_dbg_arm && c->d_dbg2 && c->gemv_dbg_slot < 1024) {
      p.dbg = c->d_dbg2;
      p.dbg_slot = c->gemv_dbg_slot++;
    }
    const size_t smem = (size_t)p.n * 8 + gemv_scratch_bytes(p, c->num_sms, 1, 512, (size_t)p.n * 8);
    void* args[] = {&p};
    return launch(c, kclass, (const void*)fn, dim3(c->num_sms), dim3(512), smem, 1, args, st);
  }
  p.dbg = nullptr;
  if (c->gemv_dbg_arm && c->d_dbg2 && c->gemv_dbg_slot < 1024) {
    p.dbg = c->d_dbg2;
    p.dbg_slot = c->gemv_dbg_slot++;
  }
  const int nb = pick_nb(c, B, p.n);
  int threads = c->opt.threads;
  if (nb >= 4) threads = 256;
  gemv_fn fn = pick_kernel(kclass, nb, threads, c->opt.f64 != 0);
  if (!fn) return fail(c, L2B_EINVAL, "no gemv kernel for nb=%d threads=%d", nb, threads);
  const size_t per = c->opt.f64 ? (size_t)p.n * 8 : (size_t)p.n * 4;
  int cps = c->opt.ctas_per_sm;
  if (threads > 256 || nb > 2 || cps < 1) cps = 1;
  if (cps * nb * per > 200 * 1024) cps = 1;
  const int grid = c->num_sms * cps;
  p.B = B;
  p.sync_wait = nullptr;
  p.sync_done = nullptr;
  if (c->sync_idx >= 0 && nb == 1 && B == 1 && c->sync_idx < c->work_cap) {
    if (c->sync_idx > 0) {
      p.sync_wait = c->d_sync + c->sync_idx - 1;
      p.sync_target = c->sync_prev_grid;
    }
    if (kclass != L2B_K_CLS) p.sync_done = c->d_sync + c->sync_idx;
    c->sync_idx++;
    c->sync_prev_grid = grid;
  }
  const size_t smem = nb * per + gemv_scratch_bytes(p, grid, nb, threads, nb * per);
  for (int b0 = 0; b0 < B; b0 += nb) {
    p.b0 = b0;
    p.nact = (B - b0) < nb ? (B - b0) : nb;
    void* args[] = {&p};
    int rc = launch(c, kclass, (const void*)fn, dim3(grid), dim3(threads), smem, 1, args, st);
    if (rc) return rc;
  }
  return 0;
}

// CTAs per (sequence, head) of the stand-alone attention kernel: as many as fit the SMs, but no more than the
// KV capacity the context was created for warrants (at least 64 time steps per CTA): splitting one head over
// a cluster costs four cluster barriers + DSMEM round trips (~5 us), which only a long context pays back.
int auto_cluster(const l2b_ctx* c, int B) {
  if (c->opt.attn_cluster > 0) return c->opt.attn_cluster;
  int cs = 8;
  while (cs > 1 && c->H * B * cs > c->num_sms) cs >>= 1;
  while (cs > 1 && (c->steps + cs - 1) / cs < 64) cs >>= 1;
  return cs;
}

// ---- batched tensor-core path --------------------------------------------------
int gemm_n_for(int cols) { return cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : 256; }

// k-split so that (tiles x splits) fills whole waves of the 148 SMs without drowning the
// step in partial-sum traffic -- and so that no accumulator sees a chain longer than
// kMaxChainKb k-blocks: tcgen05.mma adds into TMEM with truncation, the error of a logit grows
// linearly with the chain (measured on 7B: 1.1e-4 at K = 2752 on ONE accumulator, 1.5e-4 with
// 256 sequences, where w2 ran as 4 splits of 2752 -- outside the 1e-4 tolerance; chains of
// <= 1408 stay below 5e-5).  G accumulators rotate inside a work item, so a split may hold
// G * kMaxChainKb k-blocks.
constexpr int kMaxChainKb = 44;
int gemm_rotation(const l2b_ctx* c, int N) {
  if (c->opt.tc_tmem_a && N <= 128) return N == 32 ? 2 : 1;   // l2b_tc3x_tmemA_matmul_kernel
  return N == 32 ? 4 : (N == 64 ? 2 : 1);                      // l2b_tc3x_matmul_kernel
}
int pick_splits(const l2b_ctx* c, int M, int K, int B) {
  if (c->opt.tc_splits > 0) return c->opt.tc_splits < c->Smax ? c->opt.tc_splits : c->Smax;
  const int tiles = (M + kBM - 1) / kBM, kblocks = (K + kBK - 1) / kBK;
  const int cap = kMaxChainKb * gemm_rotation(c, gemm_n_for(B < 256 ? B : 256));
  double best = 1e30;
  int best_s = 0;
  for (int s = 1; s <= c->Smax && (s * 4 <= kblocks || s == 1); ++s) {
    if ((kblocks + s - 1) / s > cap && s < c->Smax && (s + 1) * 4 <= kblocks) continue;  // chain too long
    const int items = tiles * s;
    const int waves = (items + c->num_sms - 1) / c->num_sms;
    const double eff = (double)items / ((double)waves * c->num_sms);
    // weight stream at the achieved wave efficiency + partial-sum write/read + ~1 us of
    // pipeline restart per item round, all in "bytes at HBM speed"
    const double bytes = (double)M * K * 4 / eff + 2.0 * s * (double)M * B * 4 + waves * 6.0e6;
    if (bytes < best) { best = bytes; best_s = s; }
  }
  return best_s > 0 ? best_s : 1;
}

int launch_gemm(l2b_ctx* c, int kclass, const float* Wt, int M, int K, const float* Xh, const float* Xl,
                int B, int* S_out, cudaStream_t st) {
  const int S = pick_splits(c, M, K, B);
  *S_out = S;
  for (int n0 = 0; n0 < B; n0 += 256) {
    const int cols = (B - n0) < 256 ? (B - n0) : 256;
    const int N = gemm_n_for(cols);
    int rc = 0;
    GemmParams g;
    g.P = c->P;
    g.Wt = Wt; g.Xh = Xh; g.Xl = Xl; g.npad = c->Bpad;
    g.dbg = (c->dbg_arm && kclass == L2B_K_GEMM_W13) ? c->d_dbg : nullptr;
    g.M = M; g.K = K; g.S = S; g.B = B; g.n0 = n0;
    g.tiles_m = (M + kBM - 1) / kBM;
    g.kblocks = (K + kBK - 1) / kBK;
    // operand ring (weight lo tiles): 3 slots; the rest of shared memory is landing ring
    const int xbytes = 2 * N * kBK * 4;
    const int op_slot = kTileA + (N <= 128 ? 0 : xbytes), land_slot = kTileA + (N <= 128 ? xbytes : 0);
    g.dop = N <= 128 ? 3 : 2;
    g.rewrite_hi = c->opt.tc_rewrite_hi;
    g.dl = (224 * 1024 - g.dop * op_slot) / land_slot;
    if (g.dl > 12) g.dl = 12;
    const void* fn = N == 32 ? (const void*)l2b_tc3x_matmul_kernel<32>
                   : N == 64 ? (const void*)l2b_tc3x_matmul_kernel<64>
                   : N == 128 ? (const void*)l2b_tc3x_matmul_kernel<128>
                              : (const void*)l2b_tc3x_matmul_kernel<256>;
    size_t smem_bytes = (size_t)g.dl * land_slot + (size_t)g.dop * op_slot + 1024;
    int threads = kGemmThreads;
    if (c->opt.tc_tmem_a && N <= 128) {
      // weight operand in tensor memory: no operand ring in shared memory, all of it is landing ring
      fn = N == 32 ? (const void*)l2b_tc3x_tmemA_matmul_kernel<32>
         : N == 64 ? (const void*)l2b_tc3x_tmemA_matmul_kernel<64>
                   : (const void*)l2b_tc3x_tmemA_matmul_kernel<128>;
      g.dl = (224 * 1024) / land_slot;
      if (g.dl > 12) g.dl = 12;
      smem_bytes = (size_t)g.dl * land_slot + 1024;
      threads = kGemmThreadsTmemA;
    }
    int items = g.tiles_m * S;
    const int grid = items < c->num_sms ? items : c->num_sms;
    void* args[] = {&g};
    rc = launch(c, kclass, fn, dim3(grid), dim3(threads), smem_bytes, 1, args, st);
    if (rc) return rc;
  }
  return 0;
}

// Which activations / cache rows a batched pass works on: B independent sequences (decode),
// or B consecutive positions of ONE sequence (prompt prefill: every entry reads and writes
// that sequence's cache, entry i at position pos0 + i).
struct BatchView {
  float *x, *xb, *q;
  int* ctl;            // CTL header + tok[B] + pos[B]
  size_t kv_off;       // floats from the layer base to the cache of entry 0
  long long kv_stride; // floats between the caches of consecutive entries (0 for prefill)
  int cls_mode;        // 0: classifier for every entry (batched decode); 1: last entry only;
                       // 2: none (a prefill chunk that is not the last one)
};

size_t tile_major_floats(int M, int K) {
  return (size_t)((M + kBM - 1) / kBM) * ((K + kBK - 1) / kBK) * (kBM * kBK);
}

// Builds (or refreshes) the tile-major weight copy the GEMMs stream.  Not capturable: called
// before a graph capture / launch sequence starts.
int ensure_tc_weights(l2b_ctx* c) {
  if (c->Wt && !c->wt_dirty) return 0;
  const int D = c->D, F = c->F, V = c->V, L = c->L;
  if (!c->Wt) {
    c->wt_off.clear();
    size_t off = 0;
    for (int l = 0; l < L; ++l) {
      c->wt_off.push_back(off); off += tile_major_floats(3 * D, D);
      c->wt_off.push_back(off); off += tile_major_floats(D, D);
      c->wt_off.push_back(off); off += tile_major_floats(2 * F, D);
      c->wt_off.push_back(off); off += tile_major_floats(D, F);
    }
    c->wt_off.push_back(off); off += tile_major_floats(V, D);
    int rc = dev_alloc(c, &c->Wt, off, false);
    if (rc) return rc;
  }
  auto conv = [&](const float* W, size_t off, int M, int K) -> int {
    const int tiles_m = (M + kBM - 1) / kBM, kblocks = (K + kBK - 1) / kBK;
    l2b_tile_major_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(W, c->Wt + off, M, K, tiles_m, kblocks);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, L2B_ECUDA, "l2b_tile_major_kernel: %s", cudaGetErrorString(e));
    return 0;
  };
  int rc = 0;
  for (int l = 0; l < L && !rc; ++l) {
    rc = conv(c->wqkv + (size_t)l * 3 * D * D, c->wt_off[4 * l + 0], 3 * D, D);
    if (!rc) rc = conv(c->wo + (size_t)l * D * D, c->wt_off[4 * l + 1], D, D);
    if (!rc) rc = conv(c->w13 + (size_t)l * 2 * F * D, c->wt_off[4 * l + 2], 2 * F, D);
    if (!rc) rc = conv(c->w2 + (size_t)l * D * F, c->wt_off[4 * l + 3], D, F);
  }
  if (!rc) rc = conv(c->wcls, c->wt_off[4 * L], V, D);
  if (rc) return rc;
  CU(c, cudaStreamSynchronize(c->stream));
  c->wt_dirty = false;
  return 0;
}

int enqueue_step_batched(l2b_ctx* c, int B, cudaStream_t st, const BatchView& v) {
  const int D = c->D, F = c->F, hs = c->hs, H = c->H, V = c->V;
  const size_t kv_seq = (size_t)H * c->steps * hs;
  const size_t kv_layer = kv_seq * c->Bmax;
  const int* tokp = v.ctl + CTL_HDR;
  const int* posp = v.ctl + CTL_HDR + B;
  int rc, S = 1;

  const BatchView* pv = &v;
  auto resid_rms = [&](const float* P, int Sp, const float* emb, const float* rms_w) -> int {
    const BatchView& vw = *pv;
    BatVecParams v;
    memset(&v, 0, sizeof v);
    v.P = P; v.S = Sp; v.B = B; v.M = D;
    v.tok_emb = emb; v.tokp = tokp; v.x = vw.x; v.rms_w = rms_w;
    v.xh = c->XhD; v.xl = c->XlD; v.npad = c->Bpad; v.D = D;
    void* args[] = {&v};
    // cluster size: slices of >= 512 floats, at most ~512 CTAs in total, rows held in registers
    int C = 8;
    while (C > 1 && (D / C < 512 || C * B > 512)) C >>= 1;
    if (B >= 128) C = 1;
    if (C == 1)
      return launch(c, L2B_K_BATCH_EPI, (const void*)l2b_bat_resid_rms_kernel, dim3(B), dim3(256), 0, 1, args, st);
    return launch(c, L2B_K_BATCH_EPI, (const void*)l2b_bat_resid_rms_cluster_kernel, dim3(C, B), dim3(256), 0, C, args,
                  st);
  };

  rc = resid_rms(nullptr, 0, c->tok_emb, c->rms_att);  // x := embedding; rmsnorm of layer 0
  if (rc) return rc;
  const int cs = auto_cluster(c, B);
  for (int l = 0; l < c->L; ++l) {
    rc = launch_gemm(c, L2B_K_GEMM_QKV, c->Wt + c->wt_off[4 * l + 0], 3 * D, D, c->XhD, c->XlD, B, &S, st);
    if (rc) return rc;
    {
      BatQkvParams q;
      memset(&q, 0, sizeof q);
      q.P = c->P; q.S = S; q.B = B; q.D = D; q.hs = hs; q.steps = c->steps;
      q.posp = posp; q.fcr = c->fcr; q.fci = c->fci; q.q = v.q;
      q.kc = c->kc + (size_t)l * kv_layer + v.kv_off; q.vc = c->vc + (size_t)l * kv_layer + v.kv_off;
      q.kv_seq_stride = v.kv_stride;
      void* args[] = {&q};
      rc = launch(c, L2B_K_BATCH_EPI, (const void*)l2b_bat_qkv_epi_kernel, dim3((3 * D / 4 + 255) / 256, B), dim3(256), 0,
                  1, args, st);
      if (rc) return rc;
    }
    {
      AttnParams a;
      memset(&a, 0, sizeof a);
      a.q = v.q;
      a.kc = c->kc + (size_t)l * kv_layer + v.kv_off;
      a.vc = c->vc + (size_t)l * kv_layer + v.kv_off;
      a.xb = v.xb;
      a.posp = posp;
      a.H = H; a.hs = hs; a.steps = c->steps;
      a.kv_b_stride = v.kv_stride;
      a.q_stride = D; a.xb_stride = D; a.xb_off = 0;
      a.stage_bytes = kAttnStageBytes;
      a.tileT = a.stage_bytes / (hs * 4);
      a.sc_cap = ((c->steps + cs - 1) / cs + 3) & ~3;
      a.tp_size = 1;
      a.xh = c->XhD; a.xl = c->XlD; a.x_npad = c->Bpad;
      void* args[] = {&a};
      if (c->opt.attn_warp && hs <= 128 && B * H >= c->opt.attn_warp) {
        // one warp per (entry, head): throughput kernel for many independent pairs
        a.sc_cap = (c->steps + 3) & ~3;
        a.nbatch = B;
        const int wpc = kAttnWarpThreads / 32;
        rc = launch(c, L2B_K_ATTN, (const void*)l2b_attn_warp_kernel, dim3((B * H + wpc - 1) / wpc),
                    dim3(kAttnWarpThreads), (size_t)wpc * a.sc_cap * 4, 1, args, st);
      } else {
        const size_t smem = (size_t)kAttnStages * kAttnStageBytes + (size_t)a.sc_cap * 4;
        rc = launch(c, L2B_K_ATTN, (const void*)l2b_attn_decode_kernel, dim3(cs, H, B), dim3(kAttnThreads), smem, cs,
                    args, st);
      }
      if (rc) return rc;
    }
    rc = launch_gemm(c, L2B_K_GEMM_WO, c->Wt + c->wt_off[4 * l + 1], D, D, c->XhD, c->XlD, B, &S, st);
    if (rc) return rc;
    rc = resid_rms(c->P, S, nullptr, c->rms_ffn + (size_t)l * D);
    if (rc) return rc;
    rc = launch_gemm(c, L2B_K_GEMM_W13, c->Wt + c->wt_off[4 * l + 2], 2 * F, D, c->XhD, c->XlD, B, &S, st);
    if (rc) return rc;
    {
      BatSwigluParams w;
      memset(&w, 0, sizeof w);
      w.P = c->P; w.S = S; w.B = B; w.F = F; w.xh = c->XhF; w.xl = c->XlF; w.npad = c->Bpad;
      void* args[] = {&w};
      rc = launch(c, L2B_K_BATCH_EPI, (const void*)l2b_bat_swiglu_kernel, dim3((F / 4 + 255) / 256, B), dim3(256), 0, 1, args,
                  st);
      if (rc) return rc;
    }
    rc = launch_gemm(c, L2B_K_GEMM_W2, c->Wt + c->wt_off[4 * l + 3], D, F, c->XhF, c->XlF, B, &S, st);
    if (rc) return rc;
    rc = resid_rms(c->P, S, nullptr, l + 1 < c->L ? c->rms_att + (size_t)(l + 1) * D : c->rms_final);
    if (rc) return rc;
  }
  if (v.cls_mode == 2) return 0;
  if (v.cls_mode == 1) {
    // prefill: only the last position's logits exist in the reference's state afterwards
    // (llama2.ts:468-473 overwrites s.logits every step); classifier as a batch-1 matvec
    GemvParams p;
    memset(&p, 0, sizeof p);
    p.W = c->wcls; p.rows = V; p.n = D;
    p.vin = v.x + (size_t)(B - 1) * D; p.vin_stride = D;
    p.rms_w = c->rms_final;
    p.logits = c->logits; p.V = V;
    p.tokp = v.ctl + CTL_HDR; p.posp = v.ctl + CTL_HDR + B;
    p.x = v.x; p.xdim = D;
    p.ctl = c->d_pfctl; p.ticket = c->d_dev; p.next = c->d_dev + 1;
    p.forced = c->d_forced; p.out_tokens = c->d_out;
    p.blk_val = c->blk_val; p.blk_idx = c->blk_idx;
    p.evict_first = c->weight_bytes > (size_t)100 * 1024 * 1024;
    return launch_gemv(c, L2B_K_CLS, p, 1, st);
  }
  rc = launch_gemm(c, L2B_K_GEMM_CLS, c->Wt + c->wt_off[4 * c->L], V, D, c->XhD, c->XlD, B, &S, st);
  if (rc) return rc;
  {
    BatLogitsParams g;
    memset(&g, 0, sizeof g);
    g.P = c->P; g.S = S; g.B = B; g.V = V; g.logits = c->logits; g.ctl = v.ctl;
    g.next = c->d_dev + 1; g.forced = c->d_forced; g.out_tokens = c->d_out;
    void* args[] = {&g};
    rc = launch(c, L2B_K_BATCH_EPI, (const void*)l2b_bat_logits_kernel, dim3(B), dim3(1024), 0, 1, args, st);
    if (rc) return rc;
    int* ctl = v.ctl;
    void* args2[] = {&ctl};
    rc = launch(c, L2B_K_BATCH_EPI, (const void*)l2b_bat_step_kernel, dim3(1), dim3(32), 0, 1, args2, st);
    if (rc) return rc;
  }
  return 0;
}

// ---- tensor-parallel step (batch 1, row-sharded projections, in-kernel NVLink exchange) -----
int enqueue_step_tp(l2b_ctx* c, cudaStream_t st) {
  const int D = c->D, F = c->F, hs = c->hs, G = c->tp_size, R = c->tp_rank;
  const int Dl = c->Dl, Fl = c->Fl, Vl = c->Vl, Hl = c->Hl, L = c->L;
  const size_t kv_seq = (size_t)Hl * c->steps * hs;
  int ef = c->opt.evict_first;
  if (ef < 0) ef = c->weight_bytes > (size_t)100 * 1024 * 1024;

  auto flag_on = [&](int g, int e, int src) { return (int*)(c->peer[g] + c->off_flags) + (size_t)e * kMaxTp + src; };
  auto fill_tp = [&](TpParams& t, int wait_e, int out_e, size_t out_vec_off, int out_off) {
    memset(&t, 0, sizeof t);
    t.rank = R; t.size = G;
    t.epoch = c->tp_epoch; t.ticket = c->tp_ticket; t.err = c->tp_err;
    t.ll_in = wait_e >= 0 ? 1 : 0;
    t.wait_idx = wait_e < 0 ? 0 : wait_e;
    t.out_idx = out_e < 0 ? 0 : out_e;
    t.out_off = out_off;
    for (int g = 0; g < G; ++g) {
      t.peer_flags[g] = out_e >= 0 ? flag_on(g, out_e, R) : nullptr;
      t.peer_out[g] = (float*)(c->peer[g] + out_vec_off);
      t.peer_am_val[g] = (float*)(c->peer[g] + c->off_amv);
      t.peer_am_idx[g] = (int*)(c->peer[g] + c->off_ami);
    }
  };

  GemvParams base;
  memset(&base, 0, sizeof base);
  base.tokp = c->d_ctl + CTL_HDR;
  base.posp = c->d_ctl + CTL_HDR + 1;
  base.x = c->x;
  base.xdim = D;
  base.fcr = c->fcr;
  base.fci = c->fci;
  base.hs = hs;
  base.steps = c->steps;
  base.kv_seq_stride = (long long)kv_seq;
  base.ctl = c->d_ctl;
  base.ticket = c->d_dev;
  base.next = c->d_dev + 1;
  base.forced = c->d_forced;
  base.out_tokens = c->d_out;
  base.blk_val = c->blk_val;
  base.blk_idx = c->blk_idx;
  base.evict_first = ef;

  // same time-step split as the single-GPU launch, so the attention sums associate identically
  int cs = c->opt.attn_cluster;
  if (cs <= 0) {   // the rule of auto_cluster() with the MODEL's head count (not this rank's)
    cs = 8;
    while (cs > 1 && c->H * cs > c->num_sms) cs >>= 1;
    while (cs > 1 && (c->steps + cs - 1) / cs < 64) cs >>= 1;
  }
  // Fused q/k/v+attention per head under tensor parallelism: only on request (fuse_cluster > 0).  A rank
  // owns H/G heads, so one cluster per head leaves most of its SMs without weights to stream: measured on
  // 7B, 2 ranks 327 (clusters of 4) vs 355 tok/s with the stand-alone kernels (q/k/v on all 148 SMs), 4
  // ranks 483 vs 483 (before the larger attention ring), 8 ranks 607 vs 703.
  int fcs = c->opt.fuse_cluster > 0 ? c->opt.fuse_cluster : 4;
  while (fcs > 1 && (Hl * fcs > c->num_sms || 3 * hs / 2 < fcs)) fcs >>= 1;
  const bool fuse = c->opt.fuse_qkv_attn && c->opt.fuse_cluster > 0 && hs % 4 == 0 && hs <= kAttnMaxHs && c->opt.f64;
  for (int l = 0; l < L; ++l) {
    const int eA = 4 * l, eB = 4 * l + 1, eC = 4 * l + 2, eD = 4 * l + 3;
    if (fuse) {  // this rank's heads: q/k/v rows + attention in one cluster kernel per head
      QkvAttnParams f;
      memset(&f, 0, sizeof f);
      f.W = c->wqkv + (size_t)l * 3 * Dl * D;
      f.D = D; f.H = Hl; f.hs = hs; f.steps = c->steps;
      f.seg_rows = Dl;
      f.vin = c->x;
      f.rms_w = c->rms_att + (size_t)l * D;
      f.tok_emb = (l == 0) ? c->tok_emb : nullptr;
      f.tokp = c->d_ctl + CTL_HDR;
      f.posp = c->d_ctl + CTL_HDR + 1;
      f.x = c->x;
      f.q = c->q;
      f.kc = c->kc + (size_t)l * kv_seq;
      f.vc = c->vc + (size_t)l * kv_seq;
      f.fcr = c->fcr; f.fci = c->fci;
      f.xb = c->xb;
      f.stage_bytes = kAttnStageBytesBig;
      f.tileT = f.stage_bytes / (hs * 4);
      f.sc_cap = ((c->steps + fcs - 1) / fcs + 3) & ~3;
      f.evict_first = ef;
      f.l2_prefetch = 0;
      f.tp_size = G;
      f.tp_ll_in = l == 0 ? 0 : 1;
      f.tp_epoch = c->tp_epoch;
      f.tp_wait_idx = l == 0 ? 0 : eD - 4;
      f.tp_out_idx = eA;
      f.xb_off = R * Dl;
      f.tp_err = c->tp_err;
      for (int g = 0; g < G; ++g) f.peer_xb[g] = (float*)(c->peer[g] + c->off_xb);
      const size_t smem = (size_t)D * 8 + (size_t)kAttnStages * f.stage_bytes + (size_t)f.sc_cap * 4;
      void* args[] = {&f};
      int rc = launch(c, L2B_K_QKV, (const void*)l2b_qkv_attn_tp_kernel, dim3(fcs, Hl, 1), dim3(kFThreads), smem, fcs,
                      args, st);
      if (rc) return rc;
    } else {
    {  // rmsnorm -> this rank's heads of q,k,v -> RoPE -> KV write (needs x from every rank)
      GemvParams p = base;
      p.W = c->wqkv + (size_t)l * 3 * Dl * D;
      p.rows = 3 * Dl;
      p.n = D;
      p.vin = c->x;
      p.vin_stride = D;
      p.rms_w = c->rms_att + (size_t)l * D;
      p.tok_emb = (l == 0) ? c->tok_emb : nullptr;
      p.q = c->q;
      p.kc = c->kc + (size_t)l * kv_seq;
      p.vc = c->vc + (size_t)l * kv_seq;
      p.Dq = Dl;
      fill_tp(p.tp, l == 0 ? -1 : eD - 4, -1, c->off_x, 0);
      int rc = launch_gemv(c, L2B_K_QKV, p, 1, st);
      if (rc) return rc;
    }
    {  // attention over this rank's heads; output slice all-gathered into every replica of xb
      AttnParams a;
      memset(&a, 0, sizeof a);
      a.q = c->q;
      a.kc = c->kc + (size_t)l * kv_seq;
      a.vc = c->vc + (size_t)l * kv_seq;
      a.xb = c->xb;
      a.posp = c->d_ctl + CTL_HDR + 1;
      a.H = Hl;
      a.hs = hs;
      a.steps = c->steps;
      a.kv_b_stride = (long long)kv_seq;
      a.q_stride = Dl;
      a.xb_stride = D;
      a.xb_off = R * Dl;
      a.stage_bytes = kAttnStageBytesBig;
      a.tileT = a.stage_bytes / (hs * 4);
      a.sc_cap = ((c->steps + cs - 1) / cs + 3) & ~3;
      a.tp_size = G;
      a.tp_epoch = c->tp_epoch;
      a.tp_out_idx = eA;
      for (int g = 0; g < G; ++g) a.peer_xb[g] = (float*)(c->peer[g] + c->off_xb);
      const size_t smem = (size_t)kAttnStages * a.stage_bytes + (size_t)a.sc_cap * 4;
      void* args[] = {&a};
      int rc = launch(c, L2B_K_ATTN, (const void*)l2b_attn_decode_kernel, dim3(cs, Hl, 1), dim3(kAttnThreads), smem,
                      cs, args, st);
      if (rc) return rc;
    }
    }
    {  // rows [R*Dl, (R+1)*Dl) of wo + residual -> every replica of x
      GemvParams p = base;
      p.W = c->wo + (size_t)l * Dl * D;
      p.rows = Dl;
      p.n = D;
      p.vin = c->xb;
      p.vin_stride = D;
      fill_tp(p.tp, eA, eB, c->off_x, R * Dl);
      int rc = launch_gemv(c, L2B_K_WO, p, 1, st);
      if (rc) return rc;
    }
    {  // this rank's hidden units of w1/w3 -> SwiGLU -> every replica of hb
      GemvParams p = base;
      p.W = c->w13 + (size_t)l * 2 * Fl * D;
      p.rows = 2 * Fl;
      p.n = D;
      p.vin = c->x;
      p.vin_stride = D;
      p.rms_w = c->rms_ffn + (size_t)l * D;
      p.hb = c->hb;
      p.hb_stride = F;
      fill_tp(p.tp, eB, eC, c->off_hb, R * Fl);
      int rc = launch_gemv(c, L2B_K_W13, p, 1, st);
      if (rc) return rc;
    }
    {  // rows of w2 + residual -> every replica of x
      GemvParams p = base;
      p.W = c->w2 + (size_t)l * Dl * F;
      p.rows = Dl;
      p.n = F;
      p.vin = c->hb;
      p.vin_stride = F;
      fill_tp(p.tp, eC, eD, c->off_x, R * Dl);
      int rc = launch_gemv(c, L2B_K_W2, p, 1, st);
      if (rc) return rc;
    }
  }
  const int eE = 4 * L;
  {  // this rank's vocabulary rows -> every replica of logits + per-rank argmax candidate
    GemvParams p = base;
    p.W = c->wcls;
    p.rows = Vl;
    p.n = D;
    p.vin = c->x;
    p.vin_stride = D;
    p.rms_w = c->rms_final;
    p.logits = c->logits;
    p.V = c->V;
    fill_tp(p.tp, eE - 1, eE, c->off_logits, R * Vl);
    int rc = launch_gemv(c, L2B_K_CLS, p, 1, st);
    if (rc) return rc;
  }
  {
    TpFinalParams f;
    memset(&f, 0, sizeof f);
    f.size = G;
    f.n_exchanges = 4 * L + 1;
    f.epoch = c->tp_epoch;
    f.wait_flags = flag_on(R, eE, 0);
    f.wait_idx = eE;
    f.am_val = (const float*)(c->xchg + c->off_amv);
    f.am_idx = (const int*)(c->xchg + c->off_ami);
    f.logits = c->logits;
    f.ctl = c->d_ctl;
    f.next = c->d_dev + 1;
    f.forced = c->d_forced;
    f.out_tokens = c->d_out;
    f.err = c->tp_err;
    void* args[] = {&f};
    int rc = launch(c, L2B_K_CLS, (const void*)l2b_tp_finalize_kernel, dim3(1), dim3(32), 0, 1, args, st);
    if (rc) return rc;
  }
  return 0;
}

// One decode step for B sequences: tokens/positions are read from d_ctl.
int enqueue_step(l2b_ctx* c, int B, cudaStream_t st) {
  c->sync_idx = -1;
  if (c->tp_size > 1) return enqueue_step_tp(c, st);
  if (c->opt.tc_min_batch > 0 && B >= c->opt.tc_min_batch && c->P != nullptr) {
    BatchView v = {c->x, c->xb, c->q, c->d_ctl, 0, (long long)c->H * c->steps * c->hs, 0};
    return enqueue_step_batched(c, B, st, v);
  }
  const int D = c->D, F = c->F, hs = c->hs, H = c->H;
  const size_t kv_seq = (size_t)H * c->steps * hs;
  const size_t kv_layer = kv_seq * c->Bmax;
  int ef = c->opt.evict_first;
  if (ef < 0) ef = c->weight_bytes > (size_t)100 * 1024 * 1024;

  GemvParams base;
  memset(&base, 0, sizeof base);
  base.tokp = c->d_ctl + CTL_HDR;
  base.posp = c->d_ctl + CTL_HDR + B;
  base.x = c->x;
  base.xdim = D;
  base.fcr = c->fcr;
  base.fci = c->fci;
  base.hs = hs;
  base.steps = c->steps;
  base.kv_seq_stride = (long long)kv_seq;
  base.ctl = c->d_ctl;
  base.ticket = c->d_dev;
  base.next = c->d_dev + 1;
  base.forced = c->d_forced;
  base.out_tokens = c->d_out;
  base.blk_val = c->blk_val;
  base.blk_idx = c->blk_idx;
  base.evict_first = ef;
  base.l2_prefetch = ef ? c->opt.l2_prefetch : 0;  // L2-resident models need no prefetch

  const int cs = auto_cluster(c, B);
  // fused q/k/v + attention: one cluster per head
  int fcs = c->opt.fuse_cluster > 0 ? c->opt.fuse_cluster : (c->H * 8 <= 96 ? 8 : 4);
  while (fcs > 1 && (c->H * fcs > c->num_sms || 3 * hs / 2 < fcs)) fcs >>= 1;
  const size_t fuse_smem = (size_t)D * 8 + (size_t)kAttnStages * kAttnStageBytesBig +
                           (size_t)((((c->steps + fcs - 1) / fcs + 3) & ~3)) * 4;
  const bool fuse = c->opt.fuse_qkv_attn && B == 1 && hs % 4 == 0 && hs <= kAttnMaxHs && c->opt.f64 &&
                    fuse_smem <= 200 * 1024;
  c->sync_idx = -1;
  if (fuse && c->opt.soft_sync && 4 * c->L + 1 <= c->work_cap) {
    CU(c, cudaMemsetAsync(c->d_sync, 0, sizeof(int) * (size_t)c->work_cap, st));
    c->sync_idx = 0;
    c->sync_prev_grid = 0;
  }
  for (int l = 0; l < c->L; ++l) {
    if (fuse) {  // rmsnorm -> q,k,v -> RoPE -> KV write -> attention   (llama2.ts:216-267)
      QkvAttnParams f;
      memset(&f, 0, sizeof f);
      f.W = c->wqkv + (size_t)l * 3 * D * D;
      f.D = D; f.H = H; f.hs = hs; f.steps = c->steps;
      f.seg_rows = D;
      f.vin = c->x;
      f.rms_w = c->rms_att + (size_t)l * D;
      f.tok_emb = (l == 0) ? c->tok_emb : nullptr;
      f.tokp = c->d_ctl + CTL_HDR;
      f.posp = c->d_ctl + CTL_HDR + B;
      f.x = c->x;
      f.q = c->q;
      f.kc = c->kc + (size_t)l * kv_layer;
      f.vc = c->vc + (size_t)l * kv_layer;
      f.fcr = c->fcr; f.fci = c->fci;
      f.xb = c->xb;
      f.stage_bytes = kAttnStageBytesBig;
      f.tileT = f.stage_bytes / (hs * 4);
      f.sc_cap = ((c->steps + fcs - 1) / fcs + 3) & ~3;
      f.evict_first = ef;
      f.l2_prefetch = ef ? c->opt.l2_prefetch : 0;
      if (ef && c->opt.fuse_prefetch) {
        f.pf_ptr = reinterpret_cast<const unsigned char*>(c->wo + (size_t)l * D * D);
        f.pf_bytes = (long long)D * D * sizeof(float) * c->opt.fuse_prefetch / 100;
      }
      if (c->sync_idx >= 0) {
        if (c->sync_idx > 0) {
          f.sync_wait = c->d_sync + c->sync_idx - 1;
          f.sync_target = c->sync_prev_grid;
        }
        f.sync_done = c->d_sync + c->sync_idx;
        c->sync_idx++;
        c->sync_prev_grid = fcs * H;
      }
      const size_t smem = (size_t)D * 8 + (size_t)kAttnStages * f.stage_bytes + (size_t)f.sc_cap * 4;
      void* args[] = {&f};
      int rc = launch(c, L2B_K_QKV, (const void*)l2b_qkv_attn_kernel, dim3(fcs, H, 1), dim3(kFThreads), smem, fcs, args,
                      st);
      if (rc) return rc;
    } else {
    {  // rmsnorm -> q,k,v -> RoPE -> KV write   (llama2.ts:216-240)
      GemvParams p = base;
      p.W = c->wqkv + (size_t)l * 3 * D * D;
      p.rows = 3 * D;
      p.n = D;
      p.vin = c->x;
      p.vin_stride = D;
      p.rms_w = c->rms_att + (size_t)l * D;
      p.tok_emb = (l == 0) ? c->tok_emb : nullptr;
      p.q = c->q;
      p.kc = c->kc + (size_t)l * kv_layer;
      p.vc = c->vc + (size_t)l * kv_layer;
      p.Dq = D;
      int rc = launch_gemv(c, L2B_K_QKV, p, B, st);
      if (rc) return rc;
    }
    {  // attention (llama2.ts:244-267)
      AttnParams a;
      memset(&a, 0, sizeof a);
      a.q = c->q;
      a.kc = c->kc + (size_t)l * kv_layer;
      a.vc = c->vc + (size_t)l * kv_layer;
      a.xb = c->xb;
      a.posp = c->d_ctl + CTL_HDR + B;
      a.H = H;
      a.hs = hs;
      a.steps = c->steps;
      a.kv_b_stride = (long long)kv_seq;
      a.q_stride = D;
      a.xb_stride = D;
      a.xb_off = 0;
      a.stage_bytes = B == 1 ? kAttnStageBytesBig : kAttnStageBytes;
      a.tileT = a.stage_bytes / (hs * 4);
      a.sc_cap = ((c->steps + cs - 1) / cs + 3) & ~3;
      a.tp_size = 1;
      if (ef && c->opt.attn_prefetch) {
        a.pf_ptr = reinterpret_cast<const unsigned char*>(c->wo + (size_t)l * D * D);
        a.pf_bytes = (long long)D * D * sizeof(float);
      }
      const size_t smem = (size_t)kAttnStages * a.stage_bytes + (size_t)a.sc_cap * 4;
      void* args[] = {&a};
      int rc = launch(c, L2B_K_ATTN, (const void*)l2b_attn_decode_kernel, dim3(cs, H, B),
                      dim3(kAttnThreads), smem, cs, args, st);
      if (rc) return rc;
    }
    }
    {  // wo matvec + residual (llama2.ts:270-273)
      GemvParams p = base;
      p.W = c->wo + (size_t)l * D * D;
      p.rows = D;
      p.n = D;
      p.vin = c->xb;
      p.vin_stride = D;
      int rc = launch_gemv(c, L2B_K_WO, p, B, st);
      if (rc) return rc;
    }
    {  // rmsnorm -> w1,w3 -> SwiGLU (llama2.ts:276-289)
      GemvParams p = base;
      p.W = c->w13 + (size_t)l * 2 * F * D;
      p.rows = 2 * F;
      p.n = D;
      p.vin = c->x;
      p.vin_stride = D;
      p.rms_w = c->rms_ffn + (size_t)l * D;
      p.hb = c->hb;
      p.hb_stride = F;
      int rc = launch_gemv(c, L2B_K_W13, p, B, st);
      if (rc) return rc;
    }
    {  // w2 matvec + residual (llama2.ts:292-295)
      GemvParams p = base;
      p.W = c->w2 + (size_t)l * D * F;
      p.rows = D;
      p.n = F;
      p.vin = c->hb;
      p.vin_stride = F;
      int rc = launch_gemv(c, L2B_K_W2, p, B, st);
      if (rc) return rc;
    }
  }
  {  // final rmsnorm -> classifier -> argmax (llama2.ts:299-302, 364-366)
    GemvParams p = base;
    p.W = c->wcls;
    p.rows = c->V;
    p.n = D;
    p.vin = c->x;
    p.vin_stride = D;
    p.rms_w = c->rms_final;
    p.logits = c->logits;
    p.V = c->V;
    int rc = launch_gemv(c, L2B_K_CLS, p, B, st);
    c->sync_idx = -1;
    if (rc) return rc;
  }
  return 0;
}

#ifdef L2B_EXPERIMENTS
// Batch-1 persistent kernel: one cooperative launch runs n_steps decode steps.
int launch_mega(l2b_ctx* c, int n_steps, cudaStream_t st) {
  MegaParams m;
  memset(&m, 0, sizeof m);
  m.D = c->D; m.F = c->F; m.L = c->L; m.H = c->H; m.hs = c->hs; m.V = c->V; m.steps = c->steps;
  m.tok_emb = c->tok_emb; m.rms_att = c->rms_att; m.wqkv = c->wqkv; m.wo = c->wo; m.rms_ffn = c->rms_ffn;
  m.w13 = c->w13; m.w2 = c->w2; m.rms_final = c->rms_final; m.fcr = c->fcr; m.fci = c->fci; m.wcls = c->wcls;
  m.x = c->x; m.xb = c->xb; m.q = c->q; m.hb = c->hb; m.logits = c->logits; m.kc = c->kc; m.vc = c->vc;
  m.kv_layer = (long long)c->H * c->steps * c->hs * c->Bmax;
  m.ctl = c->d_ctl; m.next = c->d_dev + 1; m.forced = c->d_forced; m.out_tokens = c->d_out;
  m.blk_val = c->blk_val; m.blk_idx = c->blk_idx; m.bar = c->d_bar;
  m.n_steps = n_steps;
  int ef = c->opt.evict_first;
  if (ef < 0) ef = c->weight_bytes > (size_t)100 * 1024 * 1024;
  m.evict_first = ef;
  const int nmax = c->D > c->F ? c->D : c->F;
  size_t smem = (size_t)nmax * 8;
  const size_t attn = ((size_t)((c->steps + 3) & ~3) + (size_t)kMegaWarps * kAttnMaxHs) * 4;
  if (attn > smem) smem = attn;
  smem += (size_t)kMegaWarps * kMU * 2 * 32 * 16;  // prefetch stage
  const void* fn = (const void*)mega_decode_kernel;
  if (!c->smem_set.count(fn)) {
    cudaFuncAttributes fa;
    CU(c, cudaFuncGetAttributes(&fa, fn));
    CU(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               227 * 1024 - (int)fa.sharedSizeBytes));
    c->smem_set.insert(fn);
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(c->num_sms);
  cfg.blockDim = dim3(kMegaThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  memset(at, 0, sizeof at);
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  void* args[] = {&m};
  CU(c, cudaLaunchKernelExC(&cfg, fn, args));
  c->launch_counter++;
  return 0;
}

bool use_mega(const l2b_ctx* c, int B) {
  return c->opt.mega == 1 && B == 1 && c->Bmax == 1 && c->tp_size == 1 && !c->profiling && c->H <= c->num_sms;
}

// shared memory plan of the streaming kernel: activation vector (or attention scratch) + rings
size_t stream_act_bytes(const l2b_ctx* c) {
  const int nmax = c->D > c->F ? c->D : c->F;
  size_t act = (size_t)nmax * 8;
  const size_t attn = ((size_t)((c->steps + 3) & ~3) + (size_t)kSWarps * kAttnMaxHs + 3 * (size_t)kAttnMaxHs) * 4;
  return act > attn ? act : attn;
}
int stream_stages(const l2b_ctx* c, size_t static_smem) {
  const size_t total = 227 * 1024;
  const size_t act = stream_act_bytes(c);
  if (static_smem + act >= total) return 0;
  int s = (int)((total - static_smem - act) / ((size_t)kSWarps * kSStageFloats * 4));
  if (s > kSMaxStages) s = kSMaxStages;
  if (c->opt.stream_stages > 0 && c->opt.stream_stages < s) s = c->opt.stream_stages;
  return s;
}
bool use_stream(const l2b_ctx* c, int B) {
  if (!(c->opt.mega == 2 && B == 1 && c->tp_size == 1 && !c->profiling)) return false;
  const int nmax = c->D > c->F ? c->D : c->F;
  return c->hs <= kAttnMaxHs && c->hs % 4 == 0 && nmax <= kSMaxV4 * kSThreads * 4 && c->D <= kSMaxV4D * kSThreads * 4 && c->D % 4 == 0 && c->F % 4 == 0 &&
         2 * (c->D / 2 / c->num_sms + 2) <= kSMaxOwn && stream_stages(c, 4096) >= 2;
}

// Batch-1 streaming kernel: one cooperative launch (co-residency: the CTAs poll each other's
// outputs) runs n_steps decode steps of sequence 0.
int launch_stream(l2b_ctx* c, int n_steps, cudaStream_t st) {
  const size_t D = c->D, F = c->F, H = c->H;
  const int part_stride = c->hs + 4;
  // words: q, k, v | error flag | one mailbox per CTA {x, hb, attention partials, argmax candidates}
  const size_t mbox = (D + F + H * kSMaxChunks * (size_t)part_stride + 2 * (size_t)c->num_sms + 3) & ~(size_t)3;
  const size_t words = 3 * D + 4 + mbox * (size_t)c->num_sms;
  if (!c->d_ll) {
    CU(c, cudaMalloc((void**)&c->d_ll, words * sizeof(uint2)));
    CU(c, cudaMemsetAsync(c->d_ll, 0, words * sizeof(uint2), st));
    c->ll_words = words;
    c->ll_seq = 1;
  }
  const unsigned need = (unsigned)n_steps * (unsigned)(c->L + 1) * SK_KINDS;
  if (c->ll_seq > 0xE0000000u - need) {  // sequence numbers about to wrap: start over on clean words
    CU(c, cudaMemsetAsync(c->d_ll, 0, c->ll_words * sizeof(uint2), st));
    c->ll_seq = 1;
  }
  StreamParams m;
  memset(&m, 0, sizeof m);
  m.D = c->D; m.F = c->F; m.L = c->L; m.H = c->H; m.hs = c->hs; m.V = c->V; m.steps = c->steps;
  m.tok_emb = c->tok_emb; m.rms_att = c->rms_att; m.wqkv = c->wqkv; m.wo = c->wo; m.rms_ffn = c->rms_ffn;
  m.w13 = c->w13; m.w2 = c->w2; m.rms_final = c->rms_final; m.fcr = c->fcr; m.fci = c->fci; m.wcls = c->wcls;
  m.kc = c->kc; m.vc = c->vc;
  m.kv_layer = (long long)c->H * c->steps * c->hs * c->Bmax;
  m.logits = c->logits; m.x_out = c->x;
  m.ctl = c->d_ctl; m.next = c->d_dev + 1; m.forced = c->d_forced; m.out_tokens = c->d_out;
  uint2* w = c->d_ll;
  m.q_ll = w; w += D;
  m.k_ll = w; w += D;
  m.v_ll = w; w += D;
  m.err = reinterpret_cast<int*>(w); w += 4;
  m.x_ll = w; w += D;
  m.hb_ll = w; w += F;
  m.part_ll = w; w += H * kSMaxChunks * (size_t)part_stride;
  m.am_ll = w;
  m.mbox_stride = (long long)mbox;
  m.part_stride = part_stride;
  m.seq0 = c->ll_seq + SK_KINDS;  // layer 0 refers to "layer -1" words it never reads
  c->ll_seq += need + SK_KINDS;
  m.n_steps = n_steps;
  int ef = c->opt.evict_first;
  if (ef < 0) ef = c->weight_bytes > (size_t)100 * 1024 * 1024;
  m.evict_first = ef;
  const void* fn = (const void*)stream_decode_kernel;
  cudaFuncAttributes fa;
  CU(c, cudaFuncGetAttributes(&fa, fn));
  if (!c->smem_set.count(fn)) {
    CU(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               227 * 1024 - (int)fa.sharedSizeBytes));
    c->smem_set.insert(fn);
  }
  m.stages = stream_stages(c, fa.sharedSizeBytes);
  if (m.stages < 2) return fail(c, L2B_EINVAL, "streaming kernel: model too large for the shared-memory ring");
  int gmax = c->num_sms / c->H;
  if (gmax < 1) gmax = 1;
  if (gmax > kSMaxChunks) gmax = kSMaxChunks;
  if (c->opt.stream_chunks > 0 && c->opt.stream_chunks < gmax) gmax = c->opt.stream_chunks;
  m.gmax = gmax;
  m.dbg = c->gemv_dbg_arm ? c->d_dbg2 : nullptr;
  const size_t smem = (size_t)kSWarps * m.stages * kSStageFloats * 4 + stream_act_bytes(c);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(c->num_sms);
  cfg.blockDim = dim3(kSThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  memset(at, 0, sizeof at);
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  void* args[] = {&m};
  CU(c, cudaLaunchKernelExC(&cfg, fn, args));
  c->launch_counter++;
  c->ll_used = true;
  return 0;
}

#endif  // L2B_EXPERIMENTS

// A step sequence is issued in two phases so that ONE host thread can drive several devices
// whose kernels wait for each other (single-process tensor parallel): begin_steps() does
// everything that may synchronise or capture (tile-major weights, graph instantiation),
// issue_steps() only enqueues.
struct StepPlan {
  cudaGraphExec_t ge = nullptr;   // nullptr: launch kernel by kernel
  int per_step = 0;               // kernel launches per step
};

int begin_steps(l2b_ctx* c, int B, StepPlan* plan) {
  c->last_tc = false;
  if (c->tp_size == 1 && c->opt.tc_min_batch > 0 && B >= c->opt.tc_min_batch && c->P != nullptr) {
    int rc = ensure_tc_weights(c);  // tile-major weight copy of the tensor-core path
    if (rc) return rc;
    c->last_tc = true;
  }
  plan->ge = nullptr;
  plan->per_step = 0;
  const bool use_graph = c->opt.graph != 0 && !c->profiling;
  if (!use_graph) return 0;
  auto it = c->graphs.find(B);
  if (it == c->graphs.end()) {
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    const int64_t before = c->launch_counter;
    cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return fail(c, L2B_ECUDA, "begin capture: %s", cudaGetErrorString(e));
    int rc = enqueue_step(c, B, c->stream);
    e = cudaStreamEndCapture(c->stream, &g);
    const int per_step = (int)(c->launch_counter - before);
    c->launch_counter = before;
    if (rc) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      return rc;
    }
    if (e != cudaSuccess) return fail(c, L2B_ECUDA, "end capture: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(c, L2B_ECUDA, "graph instantiate: %s", cudaGetErrorString(e));
    c->graphs[B] = ge;
    c->graph_launches[B] = per_step;
    it = c->graphs.find(B);
  }
  plan->ge = it->second;
  plan->per_step = c->graph_launches[B];
  return 0;
}

int issue_steps(l2b_ctx* c, int B, const StepPlan& plan, int count) {
  if (plan.ge) {
    for (int s = 0; s < count; ++s) CU(c, cudaGraphLaunch(plan.ge, c->stream));
    c->launch_counter += (int64_t)count * plan.per_step;
    return 0;
  }
  for (int s = 0; s < count; ++s) {
    int rc = enqueue_step(c, B, c->stream);
    if (rc) return rc;
  }
  return 0;
}

// Runs `n_steps` decode steps for B sequences on ONE context, graph-launched when enabled.
// Events ev0/ev1 bracket the device work on the ctx stream.
int run_steps(l2b_ctx* c, int B, int n_steps) {
  const int64_t l0 = c->launch_counter;
#ifdef L2B_EXPERIMENTS
  c->last_tc = false;
  if (use_mega(c, B)) {
    CU(c, cudaEventRecord(c->ev0, c->stream));
    int rc = launch_mega(c, n_steps, c->stream);
    if (rc) return rc;
    CU(c, cudaEventRecord(c->ev1, c->stream));
    c->last_launches = c->launch_counter - l0;
    return 0;
  }
  if (use_stream(c, B)) {
    CU(c, cudaEventRecord(c->ev0, c->stream));
    int rc = launch_stream(c, n_steps, c->stream);
    if (rc) return rc;
    CU(c, cudaEventRecord(c->ev1, c->stream));
    c->last_launches = c->launch_counter - l0;
    return 0;
  }
#endif
  StepPlan plan;
  int rc = begin_steps(c, B, &plan);
  if (rc) return rc;
  CU(c, cudaEventRecord(c->ev0, c->stream));
  rc = issue_steps(c, B, plan, n_steps);
  if (rc) return rc;
  CU(c, cudaEventRecord(c->ev1, c->stream));
  c->last_launches = c->launch_counter - l0;
  return 0;
}

// ---- context groups ------------------------------------------------------------------------
// Every step-type entry point works on "parts": (member context, first global sequence, count).
// A plain context is a group of one.  All members are driven from the calling thread: inputs are
// staged everywhere, then the steps are enqueued round-robin (a tensor-parallel rank's kernels
// spin on data its peers produce, so no member may be synchronised before every member has its
// work), then results are copied out and the streams are synchronised.
struct Part {
  l2b_ctx* k;
  int b0, nb;
};

bool is_group(const l2b_ctx* c) { return !c->kids.empty(); }
bool is_tp_group(const l2b_ctx* c) { return !c->kids.empty() && c->kids[0]->tp_size > 1; }

void parts_of(l2b_ctx* c, int B, std::vector<Part>* out) {
  out->clear();
  if (!is_group(c)) {
    out->push_back({c, 0, B});
  } else if (is_tp_group(c)) {
    for (l2b_ctx* k : c->kids) out->push_back({k, 0, B});   // every rank sees the same token
  } else {
    for (size_t g = 0; g < c->kids.size(); ++g) {
      const int b0 = (int)g * c->per_kid;
      if (b0 >= B) break;
      out->push_back({c->kids[g], b0, (B - b0) < c->per_kid ? (B - b0) : c->per_kid});
    }
  }
}

// error of a member -> the group handle the caller holds
int lift(l2b_ctx* c, const l2b_ctx* k, int rc) {
  if (rc && k != c) c->err = k->err;
  return rc;
}

int run_parts(l2b_ctx* c, const std::vector<Part>& parts, int n_steps) {
  if (parts.size() == 1) {
    cudaSetDevice(parts[0].k->device);
    return lift(c, parts[0].k, run_steps(parts[0].k, parts[0].nb, n_steps));
  }
  std::vector<StepPlan> plans(parts.size());
  std::vector<int64_t> l0(parts.size());
  for (size_t i = 0; i < parts.size(); ++i) {
    l2b_ctx* k = parts[i].k;
    CU(c, cudaSetDevice(k->device));
    l0[i] = k->launch_counter;
    int rc = begin_steps(k, parts[i].nb, &plans[i]);
    if (rc) return lift(c, k, rc);
  }
  for (size_t i = 0; i < parts.size(); ++i) {
    CU(c, cudaSetDevice(parts[i].k->device));
    CU(c, cudaEventRecord(parts[i].k->ev0, parts[i].k->stream));
  }
  for (int s = 0; s < n_steps; ++s)
    for (size_t i = 0; i < parts.size(); ++i) {
      CU(c, cudaSetDevice(parts[i].k->device));
      int rc = issue_steps(parts[i].k, parts[i].nb, plans[i], 1);
      if (rc) return lift(c, parts[i].k, rc);
    }
  for (size_t i = 0; i < parts.size(); ++i) {
    l2b_ctx* k = parts[i].k;
    CU(c, cudaSetDevice(k->device));
    CU(c, cudaEventRecord(k->ev1, k->stream));
    k->last_launches = k->launch_counter - l0[i];
  }
  return 0;
}

int finish(l2b_ctx* c) {
  if (c->tp_size > 1)
    CU(c, cudaMemcpyAsync(c->h_err, c->tp_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
#ifdef L2B_EXPERIMENTS
  const bool ll = c->ll_used && c->tp_size == 1;
#else
  const bool ll = false;
#endif
  if (ll) {
    int* derr = reinterpret_cast<int*>(c->d_ll + 3 * (size_t)c->D);
    CU(c, cudaMemcpyAsync(c->h_err, derr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  }
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  if (ll) {
    c->ll_used = false;
    if (*c->h_err != 0) {
      cudaMemsetAsync(c->d_ll, 0, c->ll_words * sizeof(uint2), c->stream);
      c->ll_seq = 1;
      return fail(c, L2B_ECOMM, "streaming kernel: activation exchange timed out");
    }
  }
  if (c->tp_size > 1 && *c->h_err != 0) {
    cudaMemsetAsync(c->tp_err, 0, sizeof(int), c->stream);
    return fail(c, L2B_ECOMM, "tensor-parallel exchange timed out waiting for a peer rank");
  }
  return 0;
}

int check_ready(l2b_ctx* c) {
  if (!c) return L2B_EINVAL;
  if (is_group(c)) {
    for (l2b_ctx* k : c->kids) {
      int rc = check_ready(k);
      if (rc) return lift(c, k, rc);
    }
    return 0;
  }
  if (!l2b_weights_ready(c)) return fail(c, L2B_ESTATE, "weights not fully uploaded");
  if (c->tp_size > 1 && !c->tp_connected)
    return fail(c, L2B_ESTATE, "tensor-parallel context not connected (l2b_tp_export / l2b_tp_connect)");
  return 0;
}

// B sequences of a call against the capacity of the context (group: the global batch)
int check_batch(l2b_ctx* c, int B) {
  const int cap = is_tp_group(c) ? 1 : c->Bmax;
  if (B < 1 || B > cap) return fail(c, L2B_EINVAL, "B=%d outside [1,%d]", B, cap);
  return 0;
}

int finish_parts(l2b_ctx* c, const std::vector<Part>& parts) {
  int first = 0;
  for (const Part& p : parts) {   // synchronise every member even after a failure
    cudaSetDevice(p.k->device);
    int rc = finish(p.k);
    if (rc && !first) first = lift(c, p.k, rc);
  }
  if (is_group(c)) {
    c->last_ms = 0.f;
    c->last_launches = 0;
    for (const Part& p : parts) {
      if (p.k->last_ms > c->last_ms) c->last_ms = p.k->last_ms;   // the slowest member
      c->last_launches += p.k->last_launches;
    }
  }
  return first;
}

// Validates (token,pos) of sequence b and fills the pinned header.
int stage_inputs(l2b_ctx* c, int B, const int32_t* tokens, const int32_t* pos, int step0,
                 int use_forced, int advance, int n_steps) {
  if (B < 1 || B > c->Bmax) return fail(c, L2B_EINVAL, "B=%d outside [1,%d]", B, c->Bmax);
  if (!tokens || !pos) return fail(c, L2B_EINVAL, "null tokens/pos");
  for (int b = 0; b < B; ++b) {
    if (tokens[b] < 0 || tokens[b] >= c->V)
      return fail(c, L2B_EINVAL, "token %d of sequence %d outside [0,%d)", tokens[b], b, c->V);
    if (pos[b] < 0 || n_steps > c->steps || pos[b] > c->steps - n_steps)
      return fail(c, L2B_EINVAL, "pos %d (+%d steps) of sequence %d outside the %d cached rows",
                  pos[b], n_steps, b, c->steps);
    if (pos[b] > c->n_run[b])
      return fail(c, L2B_EORDER, "pos %d of sequence %d called before positions %d..%d were run",
                  pos[b], b, c->n_run[b], pos[b] - 1);
  }
  c->h_ctl[CTL_STEP] = step0;
  c->h_ctl[CTL_USE_FORCED] = use_forced;
  c->h_ctl[CTL_ADVANCE] = advance;
  c->h_ctl[CTL_RESERVED] = 0;
  for (int b = 0; b < B; ++b) {
    c->h_ctl[CTL_HDR + b] = tokens[b];
    c->h_ctl[CTL_HDR + B + b] = pos[b];
  }
  CU(c, cudaMemcpyAsync(c->d_ctl, c->h_ctl, sizeof(int) * (CTL_HDR + 2 * B), cudaMemcpyHostToDevice,
                        c->stream));
  return 0;
}

void mark_run(l2b_ctx* c, int B, const int32_t* pos, int n_steps) {
  for (int b = 0; b < B; ++b)
    if (pos[b] + n_steps > c->n_run[b]) c->n_run[b] = pos[b] + n_steps;
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------

L2B_API int l2b_abi_version(void) { return L2B_ABI_VERSION; }

L2B_API const char* l2b_last_error(const l2b_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

static int create_common(const int32_t hdr[7], int32_t device, int32_t max_batch, int32_t max_steps,
                         int32_t tp_rank, int32_t tp_size, l2b_ctx** out) {
  DevGuard dev_guard;
  if (!hdr || !out) return fail(nullptr, L2B_EINVAL, "null hdr/out");
  *out = nullptr;
  const int D = hdr[0], F = hdr[1], L = hdr[2], H = hdr[3];
  const int V = hdr[5] < 0 ? -hdr[5] : hdr[5], S = hdr[6];
  if (D <= 0 || F <= 0 || L <= 0 || H <= 0 || V <= 0 || S <= 0)
    return fail(nullptr, L2B_EINVAL, "non-positive config field");
  if (D % H != 0) return fail(nullptr, L2B_EINVAL, "dim %d not divisible by n_heads %d", D, H);
  const int hs = D / H;
  if (D % 4 || F % 4 || hs % 4 || (V & 1))
    return fail(nullptr, L2B_EINVAL,
                "kernels need dim, hidden_dim, head_size multiples of 4 and an even vocab");
  if (hs > kAttnMaxHs) return fail(nullptr, L2B_EINVAL, "head_size %d > %d", hs, kAttnMaxHs);
  if (max_batch < 1) return fail(nullptr, L2B_EINVAL, "max_batch must be >= 1");
  if (max_steps < 0 || max_steps > S) return fail(nullptr, L2B_EINVAL, "max_steps outside [0,seq_len]");
  if (max_steps == 0) max_steps = S;
  if (tp_size < 1 || tp_size > kMaxTp || tp_rank < 0 || tp_rank >= tp_size)
    return fail(nullptr, L2B_EINVAL, "tensor parallel rank %d of %d", tp_rank, tp_size);
  if (tp_size > 1 && (H % tp_size || F % tp_size || V % (2 * tp_size) || (D / tp_size) % 2))
    return fail(nullptr, L2B_EINVAL, "heads, hidden_dim and vocab/2 must be divisible by the tensor-parallel degree");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, L2B_ECUDA, "no CUDA device (%s); this library has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= ndev) return fail(nullptr, L2B_EINVAL, "device %d of %d", device, ndev);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, L2B_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, L2B_ECUDA, "device properties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, L2B_ECUDA, "device is sm_%d%d; this library is built for sm_100a only",
                prop.major, prop.minor);

  l2b_ctx* c = new l2b_ctx();
  c->D = D; c->F = F; c->L = L; c->H = H; c->hs = hs; c->V = V; c->S = S;
  c->shared_cls = hdr[5] > 0;
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->Bmax = max_batch;
  c->steps = max_steps;
  c->tp_rank = tp_rank;
  c->tp_size = tp_size;
  c->Hl = H / tp_size; c->Dl = c->Hl * hs; c->Fl = F / tp_size; c->Vl = V / tp_size;
  c->uploaded.assign((size_t)L2B_T_COUNT * L, 0);
  c->n_run.assign(max_batch, 0);

  int rc = 0;
#define TRY(x) if (!rc) rc = (x)
  const size_t sD = D, sF = F, sL = L, sV = V, sS = S, sB = max_batch;
  const size_t sDl = c->Dl, sFl = c->Fl, sVl = c->Vl;   // == full sizes when tp_size == 1
  TRY(dev_alloc(c, &c->tok_emb, sV * sD, false));
  TRY(dev_alloc(c, &c->rms_att, sL * sD, false));
  TRY(dev_alloc(c, &c->wqkv, sL * 3 * sDl * sD, false));
  TRY(dev_alloc(c, &c->wo, sL * sDl * sD, false));
  TRY(dev_alloc(c, &c->rms_ffn, sL * sD, false));
  TRY(dev_alloc(c, &c->w13, sL * 2 * sFl * sD, false));
  TRY(dev_alloc(c, &c->w2, sL * sDl * sF, false));
  TRY(dev_alloc(c, &c->rms_final, sD, false));
  TRY(dev_alloc(c, &c->fcr, sS * (hs / 2), false));
  TRY(dev_alloc(c, &c->fci, sS * (hs / 2), false));
  if (c->shared_cls) {
    c->wcls = c->tok_emb + (size_t)tp_rank * sVl * sD;  // llama2.ts:127 (this rank's rows)
  } else {
    TRY(dev_alloc(c, &c->wcls, sVl * sD, false));
  }
  c->weight_bytes = 4 * (sL * (4 * sDl * sD + 3 * sD * sFl + 2 * sD) + sD + sVl * sD);
  if (tp_size == 1) {
    TRY(dev_alloc(c, &c->x, sB * sD, true));
    TRY(dev_alloc(c, &c->xb, sB * sD, true));
    TRY(dev_alloc(c, &c->hb, sB * sF, true));
    TRY(dev_alloc(c, &c->logits, sB * sV, true));
  } else {
    // every vector that is gathered from all ranks lives in ONE allocation that the peers map
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const int n_ex = 4 * L + 1;
    c->off_x = 0;                               // x, xb, hb: LL replicas, 8 bytes per element
    c->off_xb = al(c->off_x + sD * 8);
    c->off_hb = al(c->off_xb + sD * 8);
    c->off_logits = al(c->off_hb + sF * 8);
    c->off_amv = al(c->off_logits + sV * 4);
    c->off_ami = al(c->off_amv + kMaxTp * 4);
    c->off_flags = al(c->off_ami + kMaxTp * 4);
    c->xchg_bytes = al(c->off_flags + (size_t)n_ex * kMaxTp * 4);
    TRY(dev_alloc(c, &c->xchg, c->xchg_bytes, true));
    if (!rc) {
      c->x = (float*)(c->xchg + c->off_x);
      c->xb = (float*)(c->xchg + c->off_xb);
      c->hb = (float*)(c->xchg + c->off_hb);
      c->logits = (float*)(c->xchg + c->off_logits);
      c->peer[tp_rank] = c->xchg;
    }
    int* tpw = nullptr;
    TRY(dev_alloc(c, &tpw, 8, true));
    c->tp_epoch = tpw;
    c->tp_ticket = tpw ? tpw + 1 : nullptr;
    c->tp_err = tpw ? tpw + 2 : nullptr;   // tpw[3]: time-out in units of 2^20 clocks, written below
  }
  TRY(dev_alloc(c, &c->q, sB * sDl, true));
  const size_t kv = sL * sB * sDl * (size_t)max_steps;
  TRY(dev_alloc(c, &c->kc, kv, true));
  TRY(dev_alloc(c, &c->vc, kv, true));
  TRY(dev_alloc(c, &c->d_ctl, CTL_HDR + 2 * sB, true));
  TRY(dev_alloc(c, &c->d_dev, 1 + sB, true));
  const size_t max_grid = (size_t)c->num_sms * 4;
  TRY(dev_alloc(c, &c->blk_val, max_grid * kMaxNB, true));
  TRY(dev_alloc(c, &c->blk_idx, max_grid * kMaxNB, true));
  if (max_batch >= 3) {
    // tensor-core path scratch: activations padded to whole 256-column groups
    c->Bpad = ((max_batch + 255) / 256) * 256;
    const size_t Mmax = (size_t)(3 * D > 2 * F ? 3 * D : 2 * F) > sV ? (size_t)(3 * D > 2 * F ? 3 * D : 2 * F) : sV;
    TRY(dev_alloc(c, &c->XhD, (size_t)c->Bpad * ((sD + 31) & ~(size_t)31), true));
    TRY(dev_alloc(c, &c->XlD, (size_t)c->Bpad * ((sD + 31) & ~(size_t)31), true));
    TRY(dev_alloc(c, &c->XhF, (size_t)c->Bpad * ((sF + 31) & ~(size_t)31), true));
    TRY(dev_alloc(c, &c->XlF, (size_t)c->Bpad * ((sF + 31) & ~(size_t)31), true));
    TRY(dev_alloc(c, &c->P, (size_t)c->Smax * sB * Mmax, false));
    c->P_floats = (size_t)c->Smax * sB * Mmax;
  }
  TRY(dev_alloc(c, &c->d_bar, 4, true));
  c->work_cap = 4 * L + 8;
  TRY(dev_alloc(c, &c->d_sync, (size_t)c->work_cap, true));
  TRY(dev_alloc(c, &c->samp_f, 4 * sV, true));
  TRY(dev_alloc(c, &c->samp_i, 2 * sV, true));
  TRY(dev_alloc(c, &c->d_forced, (size_t)max_steps * sB, true));
  TRY(dev_alloc(c, &c->d_out, (size_t)max_steps * sB, true));
#undef TRY
  if (!rc) {
    cudaError_t e2 = cudaMallocHost((void**)&c->h_ctl, sizeof(int) * (CTL_HDR + 2 * sB));
    if (e2 == cudaSuccess) e2 = cudaMallocHost((void**)&c->h_logits, sizeof(float) * sB * sV);
    if (e2 == cudaSuccess) e2 = cudaMallocHost((void**)&c->h_out, sizeof(int) * (size_t)max_steps * sB);
    if (e2 == cudaSuccess) e2 = cudaMallocHost((void**)&c->h_err, sizeof(int));
    if (e2 == cudaSuccess) e2 = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->ev0);
    if (e2 == cudaSuccess) e2 = cudaEventCreate(&c->ev1);
    if (e2 != cudaSuccess) rc = fail(c, L2B_ECUDA, "host/stream setup: %s", cudaGetErrorString(e2));
  }
  if (!rc) rc = write_tp_timeout(c);
  if (!rc) {
    // the zero-fills above ran on the NULL stream; the ctx stream is non-blocking
    cudaError_t e3 = cudaDeviceSynchronize();
    if (e3 != cudaSuccess) rc = fail(c, L2B_ECUDA, "device sync: %s", cudaGetErrorString(e3));
  }
  if (rc) {
    g_create_error = c->err;
    l2b_destroy(c);
    return rc;
  }
  *out = c;
  return L2B_OK;
}

L2B_API int l2b_create(const int32_t hdr[7], int32_t device, int32_t max_batch, int32_t max_steps,
                       l2b_ctx** out) {
  return create_common(hdr, device, max_batch, max_steps, 0, 1, out);
}

L2B_API int l2b_create_tp(const int32_t hdr[7], int32_t device, int32_t max_steps, int32_t tp_rank,
                          int32_t tp_size, l2b_ctx** out) {
  return create_common(hdr, device, 1, max_steps, tp_rank, tp_size, out);
}

// Single-process multi-GPU context ("l2b_create(hdr, n_gpus, tp_degree, max_batch,
// max_steps)"): one host thread -- the reference is one JS thread, llama2.ts:468 -- drives devices
// 0..n_gpus-1.  tp_degree == 1: the batch is partitioned over the GPUs (weights replicated, no
// collective).  tp_degree == n_gpus: ONE tensor-parallel group decoding one sequence; the ranks'
// exchange blocks are mapped into each other with cudaDeviceEnablePeerAccess (no IPC handles).
L2B_API int l2b_create_multi(const int32_t hdr[7], int32_t n_gpus, int32_t tp_degree, int32_t max_batch,
                             int32_t max_steps, l2b_ctx** out) {
  DevGuard dev_guard;
  if (!hdr || !out) return fail(nullptr, L2B_EINVAL, "null hdr/out");
  *out = nullptr;
  if (n_gpus < 1 || n_gpus > kMaxTp) return fail(nullptr, L2B_EINVAL, "n_gpus %d outside [1,%d]", n_gpus, kMaxTp);
  if (tp_degree != 1 && tp_degree != n_gpus)
    return fail(nullptr, L2B_EINVAL, "tp_degree must be 1 (batch partition) or n_gpus (one tensor-parallel group)");
  if (max_batch < 1) return fail(nullptr, L2B_EINVAL, "max_batch must be >= 1");
  const bool tp = tp_degree > 1;
  if (tp && max_batch != 1) return fail(nullptr, L2B_EINVAL, "a tensor-parallel group decodes one sequence (max_batch 1)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, L2B_ECUDA, "no CUDA device (%s); this library has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (n_gpus > ndev) return fail(nullptr, L2B_EINVAL, "n_gpus %d but only %d device(s) visible", n_gpus, ndev);

  l2b_ctx* g = new l2b_ctx();
  g->per_kid = (max_batch + n_gpus - 1) / n_gpus;
  int rc = 0;
  for (int r = 0; r < n_gpus && !rc; ++r) {
    l2b_ctx* k = nullptr;
    rc = tp ? create_common(hdr, r, 1, max_steps, r, n_gpus, &k) : create_common(hdr, r, g->per_kid, max_steps, 0, 1, &k);
    if (!rc) g->kids.push_back(k);
  }
  if (!rc && tp && n_gpus > 1) {
    for (int a = 0; a < n_gpus && !rc; ++a) {
      cudaSetDevice(a);
      for (int b = 0; b < n_gpus && !rc; ++b) {
        if (a == b) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, a, b);
        if (!can) { rc = fail(nullptr, L2B_ECOMM, "device %d cannot map device %d's memory (no NVLink/P2P)", a, b); break; }
        static bool enabled[kMaxTp][kMaxTp];   // per process: a second group must not enable it again
        if (enabled[a][b]) continue;
        e = cudaDeviceEnablePeerAccess(b, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          rc = fail(nullptr, L2B_ECOMM, "cudaDeviceEnablePeerAccess(%d -> %d): %s", a, b, cudaGetErrorString(e));
        else
          enabled[a][b] = true;
        cudaGetLastError();
      }
    }
    if (!rc)
      for (l2b_ctx* k : g->kids) {
        for (int r = 0; r < n_gpus; ++r) k->peer[r] = g->kids[r]->xchg;
        k->tp_connected = true;
      }
  }
  if (rc) {
    const std::string msg = g_create_error;
    l2b_destroy(g);
    g_create_error = msg;
    return rc;
  }
  l2b_ctx* k0 = g->kids[0];
  g->D = k0->D; g->F = k0->F; g->L = k0->L; g->H = k0->H; g->hs = k0->hs; g->V = k0->V; g->S = k0->S;
  g->shared_cls = k0->shared_cls;
  g->device = 0;
  g->steps = k0->steps;
  g->Bmax = tp ? 1 : max_batch;
  *out = g;
  return L2B_OK;
}

L2B_API void l2b_destroy(l2b_ctx* c) {
  DevGuard dev_guard;
  if (!c) return;
  if (!c->kids.empty()) {
    for (l2b_ctx* k : c->kids) {   // nobody may still be spinning on a peer that is being freed
      cudaSetDevice(k->device);
      if (k->stream) cudaStreamSynchronize(k->stream);
    }
    for (l2b_ctx* k : c->kids) l2b_destroy(k);
    delete c;
    return;
  }
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  drop_graphs(c);
  for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
  if (c->tp_size > 1) {
    for (int g = 0; g < c->tp_size; ++g)
      if (g != c->tp_rank && c->peer[g] && c->peer_ipc) cudaIpcCloseMemHandle(c->peer[g]);
    if (c->xchg) cudaFree(c->xchg);
    if (c->tp_epoch) cudaFree(c->tp_epoch);
    c->x = c->xb = c->hb = c->logits = nullptr;
  }
  if (c->h_err) cudaFreeHost(c->h_err);
  if (c->h_pfctl) cudaFreeHost(c->h_pfctl);
  if (c->d_pfctl) cudaFree(c->d_pfctl);
  float* fl[] = {c->tok_emb, c->rms_att, c->wqkv, c->wo, c->rms_ffn, c->w13, c->w2, c->rms_final,
                 c->fcr, c->fci, c->shared_cls ? nullptr : c->wcls, c->x, c->xb, c->q, c->hb,
                 c->logits, c->kc, c->vc, c->blk_val, c->XhD, c->XlD, c->XhF, c->XlF, c->P,
                 c->pf_x, c->pf_xb, c->pf_q, c->Wt};
  for (float* p : fl)
    if (p) cudaFree(p);
  int* il[] = {c->d_ctl, c->d_dev, c->blk_idx, c->d_forced, c->d_out, (int*)c->d_bar, c->d_sync, c->samp_i, (int*)c->samp_f,
               (int*)c->d_ll, (int*)c->d_dbg, (int*)c->d_dbg2};
  for (int* p : il)
    if (p) cudaFree(p);
  if (c->h_ctl) cudaFreeHost(c->h_ctl);
  if (c->h_logits) cudaFreeHost(c->h_logits);
  if (c->h_out) cudaFreeHost(c->h_out);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// Where a tensor of the checkpoint lives on this context.  Row-sharded tensors keep rows
// [rank*slice, (rank+1)*slice) of the FULL tensor (slice == all rows when tp_size == 1):
// `src_off` floats of the full tensor are skipped, then `rows` rows of `row` floats follow.
struct Placement {
  float* dst;
  size_t expect;     // floats of the full tensor (what the caller / the file holds)
  size_t src_off;    // first float kept
  size_t rows, row;  // kept rows x floats per row
  size_t dst_pitch;  // floats between kept rows on the device (2*row for the w1/w3 interleave)
};

static int placement(l2b_ctx* c, int tensor_id, int layer, Placement* pl) {
  if (tensor_id < 0 || tensor_id >= L2B_T_COUNT) return fail(c, L2B_EINVAL, "tensor id %d", tensor_id);
  const size_t D = c->D, F = c->F, V = c->V, S = c->S, hs2 = c->hs / 2;
  const bool layered = (tensor_id >= L2B_T_RMS_ATT_WEIGHT && tensor_id <= L2B_T_W3);
  if (layer < 0 || layer >= (layered ? c->L : 1))
    return fail(c, L2B_EINVAL, "layer %d out of range for tensor %d", layer, tensor_id);
  const size_t Dl = c->Dl, Fl = c->Fl, Vl = c->Vl, r = c->tp_rank;
  Placement p = {nullptr, 0, 0, 1, 0, 0};
  switch (tensor_id) {
    case L2B_T_TOKEN_EMBEDDING_TABLE: p.expect = V * D; p.rows = V; p.row = D; p.dst = c->tok_emb; break;
    case L2B_T_RMS_ATT_WEIGHT: p.expect = D; p.row = D; p.dst = c->rms_att + layer * D; break;
    case L2B_T_WQ:
    case L2B_T_WK:
    case L2B_T_WV:
      p.expect = D * D; p.src_off = r * Dl * D; p.rows = Dl; p.row = D;
      p.dst = c->wqkv + (size_t)layer * 3 * Dl * D + (size_t)(tensor_id - L2B_T_WQ) * Dl * D;
      break;
    case L2B_T_WO:
      p.expect = D * D; p.src_off = r * Dl * D; p.rows = Dl; p.row = D; p.dst = c->wo + (size_t)layer * Dl * D;
      break;
    case L2B_T_RMS_FFN_WEIGHT: p.expect = D; p.row = D; p.dst = c->rms_ffn + layer * D; break;
    case L2B_T_W1:
    case L2B_T_W3:  // interleave rows: device row 2i = w1 row i, 2i+1 = w3 row i (pairs feed SwiGLU)
      p.expect = F * D; p.src_off = r * Fl * D; p.rows = Fl; p.row = D; p.dst_pitch = 2 * D;
      p.dst = c->w13 + (size_t)layer * 2 * Fl * D + (tensor_id == L2B_T_W3 ? D : 0);
      break;
    case L2B_T_W2:
      p.expect = D * F; p.src_off = r * Dl * F; p.rows = Dl; p.row = F; p.dst = c->w2 + (size_t)layer * Dl * F;
      break;
    case L2B_T_RMS_FINAL_WEIGHT: p.expect = D; p.row = D; p.dst = c->rms_final; break;
    case L2B_T_FREQ_CIS_REAL: p.expect = S * hs2; p.rows = S; p.row = hs2; p.dst = c->fcr; break;
    case L2B_T_FREQ_CIS_IMAG: p.expect = S * hs2; p.rows = S; p.row = hs2; p.dst = c->fci; break;
    case L2B_T_WCLS:
      if (c->shared_cls)
        return fail(c, L2B_ESTATE, "shared classifier: wcls aliases the embedding table (llama2.ts:127)");
      p.expect = V * D; p.src_off = r * Vl * D; p.rows = Vl; p.row = D; p.dst = c->wcls;
      break;
  }
  if (p.dst_pitch == 0) p.dst_pitch = p.row;
  *pl = p;
  return 0;
}

L2B_API int l2b_upload(l2b_ctx* c, int32_t tensor_id, int32_t layer, const float* host,
                       uint64_t n_floats) {
  DevGuard dev_guard;
  if (!c) return L2B_EINVAL;
  if (!host) return fail(c, L2B_EINVAL, "null host pointer");
  if (is_group(c)) {   // replicas keep the whole tensor, tensor-parallel ranks their rows
    for (l2b_ctx* k : c->kids) {
      int rc = l2b_upload(k, tensor_id, layer, host, n_floats);
      if (rc) return lift(c, k, rc);
    }
    return L2B_OK;
  }
  Placement pl;
  int rc = placement(c, tensor_id, layer, &pl);
  if (rc) return rc;
  if (n_floats != pl.expect)
    return fail(c, L2B_EINVAL, "tensor %d expects %zu floats, got %llu", tensor_id, pl.expect,
                (unsigned long long)n_floats);
  CU(c, cudaSetDevice(c->device));
  // stream-ordered behind any step still reading the old contents
  CU(c, cudaMemcpy2DAsync(pl.dst, pl.dst_pitch * sizeof(float), host + pl.src_off, pl.row * sizeof(float),
                          pl.row * sizeof(float), pl.rows, cudaMemcpyDefault, c->stream));
  // the copy runs on the ctx stream (device sources are asynchronous otherwise): the
  // caller may free `host` on return and the next step sees the new contents
  CU(c, cudaStreamSynchronize(c->stream));
  c->uploaded[(size_t)tensor_id * c->L + layer] = 1;
  c->wt_dirty = true;
  return L2B_OK;
}

// Checkpoint loader fast path.  The reference reads the
// llama2.c legacy-v0 file tensor by tensor into fresh Buffers (fs.readSync, llama2.ts:44-68,
// 112-129); here the file streams through two pinned staging buffers, the H2D copy of one
// overlapping the pread of the other, straight into the device layout (w1/w3 interleave,
// q/k/v stacking).  Shard-aware: a tensor-parallel rank preads only its own rows.
L2B_API int l2b_load_checkpoint(l2b_ctx* c, const char* path, double* seconds_out) {
  DevGuard dev_guard;
  if (!c || !path) return L2B_EINVAL;
  if (is_group(c)) {   // every member streams the file (tensor-parallel ranks: only their rows)
    const auto g0 = std::chrono::steady_clock::now();
    for (l2b_ctx* k : c->kids) {
      int rc = l2b_load_checkpoint(k, path, nullptr);
      if (rc) return lift(c, k, rc);
    }
    if (seconds_out) *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - g0).count();
    return L2B_OK;
  }
  CU(c, cudaSetDevice(c->device));
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(c, L2B_EINVAL, "cannot open %s", path);
  int32_t hdr[7];
  if (pread(fd, hdr, sizeof hdr, 0) != (ssize_t)sizeof hdr) {
    close(fd);
    return fail(c, L2B_EINVAL, "%s: short header", path);
  }
  const int V = hdr[5] < 0 ? -hdr[5] : hdr[5];
  if (hdr[0] != c->D || hdr[1] != c->F || hdr[2] != c->L || hdr[3] != c->H || V != c->V || hdr[6] != c->S ||
      (hdr[5] > 0) != c->shared_cls) {
    close(fd);
    return fail(c, L2B_EINVAL, "%s: header does not match the context's configuration", path);
  }
  const size_t kChunk = (size_t)32 << 20;
  unsigned char* stage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  auto cleanup = [&]() {   // every exit below goes through here
    for (int i = 0; i < 2; ++i) {
      if (stage[i]) cudaFreeHost(stage[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
    }
    close(fd);
  };
  cudaError_t e = cudaMallocHost((void**)&stage[0], kChunk);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&stage[1], kChunk);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cleanup();
    return fail(c, L2B_ENOMEM, "pinned staging: %s", cudaGetErrorString(e));
  }
  e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    cleanup();
    return fail(c, L2B_ECUDA, "stream sync: %s", cudaGetErrorString(e));
  }
  const auto t0 = std::chrono::steady_clock::now();
  // file order (llama2.ts:114-127)
  static const int order[] = {L2B_T_TOKEN_EMBEDDING_TABLE, L2B_T_RMS_ATT_WEIGHT, L2B_T_WQ, L2B_T_WK, L2B_T_WV,
                              L2B_T_WO, L2B_T_RMS_FFN_WEIGHT, L2B_T_W1, L2B_T_W2, L2B_T_W3,
                              L2B_T_RMS_FINAL_WEIGHT, L2B_T_FREQ_CIS_REAL, L2B_T_FREQ_CIS_IMAG, L2B_T_WCLS};
  size_t file_off = sizeof hdr;
  int rc = 0, slot = 0;
  bool used[2] = {false, false};
  for (int t : order) {
    if (t == L2B_T_WCLS && c->shared_cls) continue;
    const bool layered = (t >= L2B_T_RMS_ATT_WEIGHT && t <= L2B_T_W3);
    for (int l = 0; l < (layered ? c->L : 1) && !rc; ++l) {
      Placement pl;
      rc = placement(c, t, l, &pl);
      if (rc) break;
      const size_t row_bytes = pl.row * sizeof(float);
      size_t rows_per_chunk = kChunk / row_bytes;
      if (rows_per_chunk == 0) { rc = fail(c, L2B_EINVAL, "row larger than the staging buffer"); break; }
      for (size_t r0 = 0; r0 < pl.rows && !rc; r0 += rows_per_chunk) {
        const size_t nr = (pl.rows - r0) < rows_per_chunk ? (pl.rows - r0) : rows_per_chunk;
        if (used[slot]) cudaEventSynchronize(ev[slot]);  // the previous copy out of this buffer is done
        const size_t want = nr * row_bytes;
        size_t got = 0;
        const off_t at = (off_t)(file_off + (pl.src_off + r0 * pl.row) * sizeof(float));
        while (got < want) {
          const ssize_t k = pread(fd, stage[slot] + got, want - got, at + (off_t)got);
          if (k <= 0) { rc = fail(c, L2B_EINVAL, "%s: truncated at tensor %d layer %d", path, t, l); break; }
          got += (size_t)k;
        }
        if (rc) break;
        cudaError_t ce = cudaMemcpy2DAsync(pl.dst + r0 * pl.dst_pitch, pl.dst_pitch * sizeof(float), stage[slot],
                                           row_bytes, row_bytes, nr, cudaMemcpyHostToDevice, c->stream);
        if (ce == cudaSuccess) ce = cudaEventRecord(ev[slot], c->stream);
        if (ce != cudaSuccess) { rc = fail(c, L2B_ECUDA, "H2D copy: %s", cudaGetErrorString(ce)); break; }
        used[slot] = true;
        slot ^= 1;
      }
      if (!rc) { c->uploaded[(size_t)t * c->L + l] = 1; c->wt_dirty = true; }
      file_off += pl.expect * sizeof(float);
    }
    if (rc) break;
  }
  cudaStreamSynchronize(c->stream);
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (seconds_out) *seconds_out = secs;
  cleanup();
  return rc;
}

L2B_API int l2b_weights_ready(const l2b_ctx* c) {
  if (!c) return 0;
  if (!c->kids.empty()) {
    for (const l2b_ctx* k : c->kids)
      if (!l2b_weights_ready(k)) return 0;
    return 1;
  }
  for (int t = 0; t < L2B_T_COUNT; ++t) {
    const bool layered = (t >= L2B_T_RMS_ATT_WEIGHT && t <= L2B_T_W3);
    if (t == L2B_T_WCLS && c->shared_cls) continue;
    const int n = layered ? c->L : 1;
    for (int l = 0; l < n; ++l)
      if (!c->uploaded[(size_t)t * c->L + l]) return 0;
  }
  return 1;
}

L2B_API int l2b_forward_batch(l2b_ctx* c, int32_t B, const int32_t* tokens, const int32_t* pos,
                              float* logits_out, int32_t* argmax_out) {
  DevGuard dev_guard;
  int rc = check_ready(c);
  if (rc) return rc;
  if ((rc = check_batch(c, B))) return rc;
  if (!tokens || !pos) return fail(c, L2B_EINVAL, "null tokens/pos");
  std::vector<Part> parts;
  parts_of(c, B, &parts);
  for (const Part& p : parts) {
    CU(c, cudaSetDevice(p.k->device));
    rc = stage_inputs(p.k, p.nb, tokens + p.b0, pos + p.b0, 0, 0, 0, 1);
    if (rc) return lift(c, p.k, rc);
  }
  rc = run_parts(c, parts, 1);
  if (rc) return rc;
  const size_t n_out = is_tp_group(c) ? 1 : parts.size();   // every rank of a tp group holds the result
  for (size_t i = 0; i < n_out; ++i) {
    l2b_ctx* k = parts[i].k;
    CU(c, cudaSetDevice(k->device));
    if (logits_out)
      CU(c, cudaMemcpyAsync(k->h_logits, k->logits, sizeof(float) * (size_t)parts[i].nb * k->V,
                            cudaMemcpyDeviceToHost, k->stream));
    if (argmax_out)
      CU(c, cudaMemcpyAsync(k->h_out, k->d_dev + 1, sizeof(int) * parts[i].nb, cudaMemcpyDeviceToHost, k->stream));
  }
  rc = finish_parts(c, parts);
  if (rc) return rc;
  for (size_t i = 0; i < n_out; ++i) {
    const Part& p = parts[i];
    if (logits_out) memcpy(logits_out + (size_t)p.b0 * c->V, p.k->h_logits, sizeof(float) * (size_t)p.nb * c->V);
    if (argmax_out) memcpy(argmax_out + p.b0, p.k->h_out, sizeof(int) * p.nb);
  }
  for (const Part& p : parts) mark_run(p.k, p.nb, pos + p.b0, 1);
  return L2B_OK;
}

L2B_API int l2b_forward(l2b_ctx* c, int32_t token, int32_t pos, float* logits_out) {
  if (c && !logits_out) return fail(c, L2B_EINVAL, "null logits_out");
  return l2b_forward_batch(c, 1, &token, &pos, logits_out, nullptr);
}

L2B_API int l2b_forward_argmax(l2b_ctx* c, int32_t token, int32_t pos, int32_t* next_out) {
  if (c && !next_out) return fail(c, L2B_EINVAL, "null next_out");
  return l2b_forward_batch(c, 1, &token, &pos, nullptr, next_out);
}

// Scratch for l2b_prefill: activations of up to `cap` prompt positions.
static int ensure_prefill(l2b_ctx* c, int cap) {
  const size_t D = c->D, F = c->F, V = c->V;
  const size_t Mmax = (3 * D > 2 * F ? 3 * D : 2 * F) > V ? (3 * D > 2 * F ? 3 * D : 2 * F) : V;
  int rc = 0;
  if (cap > c->pf_cap) {
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->pf_x) { cudaFree(c->pf_x); cudaFree(c->pf_xb); cudaFree(c->pf_q); cudaFree(c->d_pfctl); cudaFreeHost(c->h_pfctl); }
    c->pf_x = c->pf_xb = c->pf_q = nullptr; c->d_pfctl = nullptr; c->h_pfctl = nullptr; c->pf_cap = 0;
    if (!rc) rc = dev_alloc(c, &c->pf_x, (size_t)cap * D, true);
    if (!rc) rc = dev_alloc(c, &c->pf_xb, (size_t)cap * D, true);
    if (!rc) rc = dev_alloc(c, &c->pf_q, (size_t)cap * D, true);
    if (!rc) rc = dev_alloc(c, &c->d_pfctl, CTL_HDR + 2 * (size_t)cap, true);
    if (!rc && cudaMallocHost((void**)&c->h_pfctl, sizeof(int) * (CTL_HDR + 2 * (size_t)cap)) != cudaSuccess)
      rc = fail(c, L2B_ENOMEM, "pinned prefill header");
    if (rc) return rc;
    c->pf_cap = cap;
  }
  if (!c->XhD) {
    c->Bpad = 256;
    if (!rc) rc = dev_alloc(c, &c->XhD, (size_t)c->Bpad * ((D + 31) & ~(size_t)31), true);
    if (!rc) rc = dev_alloc(c, &c->XlD, (size_t)c->Bpad * ((D + 31) & ~(size_t)31), true);
    if (!rc) rc = dev_alloc(c, &c->XhF, (size_t)c->Bpad * ((F + 31) & ~(size_t)31), true);
    if (!rc) rc = dev_alloc(c, &c->XlF, (size_t)c->Bpad * ((F + 31) & ~(size_t)31), true);
    if (rc) return rc;
  }
  const size_t need = (size_t)c->Smax * cap * Mmax;
  if (need > c->P_floats) {
    CU(c, cudaStreamSynchronize(c->stream));
    drop_graphs(c);  // graphs captured the old pointer
    if (c->P) cudaFree(c->P);
    c->P = nullptr; c->P_floats = 0;
    rc = dev_alloc(c, &c->P, need, false);
    if (rc) return rc;
    c->P_floats = need;
  }
  CU(c, cudaDeviceSynchronize());
  return 0;
}

L2B_API int l2b_prefill(l2b_ctx* c, int32_t seq, int32_t n_tokens, const int32_t* tokens, int32_t pos0,
                        float* logits_out, int32_t* argmax_out) {
  DevGuard dev_guard;
  int rc = check_ready(c);
  if (rc) return rc;
  if (c->tp_size > 1 || is_tp_group(c))
    return fail(c, L2B_ESTATE, "prefill is not available on a tensor-parallel context");
  if (seq < 0 || seq >= c->Bmax) return fail(c, L2B_EINVAL, "seq %d outside [0,%d)", seq, c->Bmax);
  if (is_group(c)) {   // the member that owns this sequence
    l2b_ctx* k = c->kids[seq / c->per_kid];
    rc = lift(c, k, l2b_prefill(k, seq % c->per_kid, n_tokens, tokens, pos0, logits_out, argmax_out));
    c->last_ms = k->last_ms;
    c->last_launches = k->last_launches;
    return rc;
  }
  if (n_tokens < 1 || !tokens) return fail(c, L2B_EINVAL, "n_tokens < 1 or null tokens");
  if (pos0 < 0 || n_tokens > c->steps || pos0 > c->steps - n_tokens)
    return fail(c, L2B_EINVAL, "positions %d..%d outside the %d cached rows", pos0, pos0 + n_tokens - 1, c->steps);
  if (pos0 > c->n_run[seq])
    return fail(c, L2B_EORDER, "pos %d of sequence %d called before positions %d..%d were run", pos0, seq,
                c->n_run[seq], pos0 - 1);
  for (int i = 0; i < n_tokens; ++i)
    if (tokens[i] < 0 || tokens[i] >= c->V) return fail(c, L2B_EINVAL, "token %d outside [0,%d)", tokens[i], c->V);
  CU(c, cudaSetDevice(c->device));
  const int cap = n_tokens < 256 ? (n_tokens < 32 ? 32 : n_tokens) : 256;
  rc = ensure_prefill(c, cap);
  if (rc) return rc;
  rc = ensure_tc_weights(c);
  if (rc) return rc;
  c->last_tc = true;
  const size_t kv_seq = (size_t)c->H * c->steps * c->hs;
  const int64_t l0 = c->launch_counter;
  CU(c, cudaEventRecord(c->ev0, c->stream));
  for (int done = 0; done < n_tokens; done += c->pf_cap) {
    const int B = (n_tokens - done) < c->pf_cap ? (n_tokens - done) : c->pf_cap;
    const bool last = done + B == n_tokens;
    if (done > 0) CU(c, cudaStreamSynchronize(c->stream));  // h_pfctl is reused
    c->h_pfctl[CTL_STEP] = 0;
    c->h_pfctl[CTL_USE_FORCED] = 0;
    c->h_pfctl[CTL_ADVANCE] = 0;
    c->h_pfctl[CTL_RESERVED] = 0;
    for (int i = 0; i < B; ++i) {
      c->h_pfctl[CTL_HDR + i] = tokens[done + i];
      c->h_pfctl[CTL_HDR + B + i] = pos0 + done + i;
    }
    CU(c, cudaMemcpyAsync(c->d_pfctl, c->h_pfctl, sizeof(int) * (CTL_HDR + 2 * B), cudaMemcpyHostToDevice,
                          c->stream));
    BatchView v = {c->pf_x, c->pf_xb, c->pf_q, c->d_pfctl, (size_t)seq * kv_seq, 0, last ? 1 : 2};
    rc = enqueue_step_batched(c, B, c->stream, v);
    if (rc) return rc;
  }
  CU(c, cudaEventRecord(c->ev1, c->stream));
  c->last_launches = c->launch_counter - l0;
  if (logits_out)
    CU(c, cudaMemcpyAsync(c->h_logits, c->logits, sizeof(float) * c->V, cudaMemcpyDeviceToHost, c->stream));
  if (argmax_out)
    CU(c, cudaMemcpyAsync(c->h_out, c->d_dev + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  rc = finish(c);
  if (rc) return rc;
  if (logits_out) memcpy(logits_out, c->h_logits, sizeof(float) * c->V);
  if (argmax_out) *argmax_out = c->h_out[0];
  if (pos0 + n_tokens > c->n_run[seq]) c->n_run[seq] = pos0 + n_tokens;
  return L2B_OK;
}

// Device sampler on `logits` (device pointer, vocab floats); result lands in d_dev[1].
static int enqueue_sampler(l2b_ctx* c, const float* logits, double temperature, double topp, double rand01) {
  SampleParams sp;
  memset(&sp, 0, sizeof sp);
  sp.logits = logits; sp.V = c->V;
  sp.temperature = temperature; sp.topp = topp; sp.rand01 = rand01;
  sp.probs = c->samp_f; sp.cand_p = c->samp_f + c->V; sp.sort_p = c->samp_f + 2 * (size_t)c->V;
  sp.cand_i = c->samp_i; sp.sort_i = c->samp_i + c->V;
  sp.next = c->d_dev + 1;
  void* args[] = {&sp};
  return launch(c, L2B_K_CLS, (const void*)l2b_sample_kernel, dim3(1), dim3(kSampThreads),
                16 * kSampThreads * sizeof(int), 1, args, c->stream);
}

L2B_API int l2b_forward_sample(l2b_ctx* c, int32_t token, int32_t pos, double temperature, double topp,
                               float rand01, int32_t* next_out) {
  DevGuard dev_guard;
  int rc = check_ready(c);
  if (rc) return rc;
  if (!next_out) return fail(c, L2B_EINVAL, "null next_out");
  if (temperature == 0.0) return l2b_forward_argmax(c, token, pos, next_out);  // llama2.ts:476-478
  std::vector<Part> parts;
  parts_of(c, 1, &parts);
  for (const Part& p : parts) {
    CU(c, cudaSetDevice(p.k->device));
    rc = stage_inputs(p.k, 1, &token, &pos, 0, 0, 0, 1);
    if (rc) return lift(c, p.k, rc);
  }
  rc = run_parts(c, parts, 1);
  if (rc) return rc;
  l2b_ctx* k = parts[0].k;   // sequence 0 / rank 0 (every tensor-parallel rank holds the full logits)
  CU(c, cudaSetDevice(k->device));
  rc = enqueue_sampler(k, k->logits, temperature, topp, (double)rand01);
  if (rc) return lift(c, k, rc);
  CU(c, cudaEventRecord(k->ev1, k->stream));
  CU(c, cudaMemcpyAsync(k->h_out, k->d_dev + 1, sizeof(int), cudaMemcpyDeviceToHost, k->stream));
  rc = finish_parts(c, parts);
  if (rc) return rc;
  *next_out = k->h_out[0];
  for (const Part& p : parts) mark_run(p.k, 1, &pos, 1);
  return L2B_OK;
}

L2B_API int l2b_sample_logits(l2b_ctx* c, const float* logits_host, double temperature, double topp, float rand01,
                              int32_t* next_out) {
  DevGuard dev_guard;
  if (!c) return L2B_EINVAL;
  if (is_group(c)) return lift(c, c->kids[0], l2b_sample_logits(c->kids[0], logits_host, temperature, topp, rand01, next_out));
  if (!logits_host || !next_out) return fail(c, L2B_EINVAL, "null argument");
  if (temperature == 0.0) return fail(c, L2B_EINVAL, "temperature 0 is the argmax path (l2b_forward_argmax)");
  CU(c, cudaSetDevice(c->device));
  float* dl = c->samp_f + 3 * (size_t)c->V;
  CU(c, cudaMemcpyAsync(dl, logits_host, sizeof(float) * c->V, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaEventRecord(c->ev0, c->stream));
  int rc = enqueue_sampler(c, dl, temperature, topp, (double)rand01);
  if (rc) return rc;
  CU(c, cudaEventRecord(c->ev1, c->stream));
  CU(c, cudaMemcpyAsync(c->h_out, c->d_dev + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  rc = finish(c);
  if (rc) return rc;
  *next_out = c->h_out[0];
  return L2B_OK;
}

L2B_API int l2b_generate_greedy(l2b_ctx* c, int32_t B, const int32_t* tokens, const int32_t* pos,
                                int32_t n_steps, const int32_t* forced, int32_t* out_tokens) {
  DevGuard dev_guard;
  int rc = check_ready(c);
  if (rc) return rc;
  if ((rc = check_batch(c, B))) return rc;
  if (!tokens || !pos) return fail(c, L2B_EINVAL, "null tokens/pos");
  if (n_steps < 1 || !out_tokens) return fail(c, L2B_EINVAL, "n_steps < 1 or null out_tokens");
  if (n_steps > c->steps) return fail(c, L2B_EINVAL, "n_steps %d > the %d cached rows", n_steps, c->steps);
  const size_t n = (size_t)n_steps * B;
  if (forced)
    for (size_t i = 0; i < n; ++i)
      if (forced[i] >= c->V) return fail(c, L2B_EINVAL, "forced token %d >= vocab", forced[i]);
  std::vector<Part> parts;
  parts_of(c, B, &parts);
  for (const Part& p : parts) {
    l2b_ctx* k = p.k;
    CU(c, cudaSetDevice(k->device));
    rc = stage_inputs(k, p.nb, tokens + p.b0, pos + p.b0, 0, forced != nullptr, 1, n_steps);
    if (rc) return lift(c, k, rc);
    if (forced) {   // this member's columns of the [n_steps][B] matrix
      for (int s = 0; s < n_steps; ++s)
        memcpy(k->h_out + (size_t)s * p.nb, forced + (size_t)s * B + p.b0, sizeof(int) * p.nb);
      CU(c, cudaMemcpyAsync(k->d_forced, k->h_out, sizeof(int) * (size_t)n_steps * p.nb, cudaMemcpyHostToDevice,
                            k->stream));
    }
  }
  if (forced)   // h_out is reused for the result below
    for (const Part& p : parts) {
      CU(c, cudaSetDevice(p.k->device));
      CU(c, cudaStreamSynchronize(p.k->stream));
    }
  rc = run_parts(c, parts, n_steps);
  if (rc) return rc;
  const size_t n_out = is_tp_group(c) ? 1 : parts.size();
  for (size_t i = 0; i < n_out; ++i) {
    l2b_ctx* k = parts[i].k;
    CU(c, cudaSetDevice(k->device));
    CU(c, cudaMemcpyAsync(k->h_out, k->d_out, sizeof(int) * (size_t)n_steps * parts[i].nb, cudaMemcpyDeviceToHost,
                          k->stream));
  }
  rc = finish_parts(c, parts);
  if (rc) return rc;
  for (size_t i = 0; i < n_out; ++i) {
    const Part& p = parts[i];
    for (int s = 0; s < n_steps; ++s)
      memcpy(out_tokens + (size_t)s * B + p.b0, p.k->h_out + (size_t)s * p.nb, sizeof(int) * p.nb);
  }
  for (const Part& p : parts) mark_run(p.k, p.nb, pos + p.b0, n_steps);
  return L2B_OK;
}

L2B_API float l2b_last_device_ms(const l2b_ctx* c) { return c ? c->last_ms : 0.f; }
L2B_API int64_t l2b_last_launches(const l2b_ctx* c) { return c ? c->last_launches : 0; }

L2B_API int l2b_profile_batch(l2b_ctx* c, int32_t B, const int32_t* tokens, const int32_t* pos,
                              float* ms_per_class, int32_t* launches_per_class) {
  DevGuard dev_guard;
  int rc = check_ready(c);
  if (rc) return rc;
  if ((rc = check_batch(c, B))) return rc;
  if (!tokens || !pos) return fail(c, L2B_EINVAL, "null tokens/pos");
  if (!ms_per_class || !launches_per_class) return fail(c, L2B_EINVAL, "null output");
  std::vector<Part> parts;
  parts_of(c, B, &parts);
  for (const Part& p : parts) {
    CU(c, cudaSetDevice(p.k->device));
    rc = stage_inputs(p.k, p.nb, tokens + p.b0, pos + p.b0, 0, 0, 0, 1);
    if (rc) return lift(c, p.k, rc);
    p.k->profiling = true;
    p.k->prof_events.clear();
    p.k->prof_class.clear();
  }
  rc = run_parts(c, parts, 1);
  for (const Part& p : parts) p.k->profiling = false;
  if (!rc) {
    for (const Part& p : parts) {
      cudaSetDevice(p.k->device);
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, p.k->stream);
      p.k->prof_events.push_back(e);
    }
    rc = finish_parts(c, parts);
  }
  for (int k = 0; k < L2B_K_COUNT; ++k) {
    ms_per_class[k] = 0.f;
    launches_per_class[k] = 0;
  }
  if (!rc) {
    l2b_ctx* k0 = parts[0].k;   // the first member's kernels (members run the same launch sequence)
    cudaSetDevice(k0->device);
    for (size_t i = 0; i + 1 < k0->prof_events.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, k0->prof_events[i], k0->prof_events[i + 1]);
      ms_per_class[k0->prof_class[i]] += ms;
      launches_per_class[k0->prof_class[i]] += 1;
    }
    for (const Part& p : parts) mark_run(p.k, p.nb, pos + p.b0, 1);
  }
  for (const Part& p : parts) {
    for (cudaEvent_t e : p.k->prof_events) cudaEventDestroy(e);
    p.k->prof_events.clear();
    p.k->prof_class.clear();
  }
  return rc;
}

L2B_API int l2b_profile_step(l2b_ctx* c, int32_t token, int32_t pos, float* ms_per_class,
                             int32_t* launches_per_class) {
  return l2b_profile_batch(c, 1, &token, &pos, ms_per_class, launches_per_class);
}

L2B_API int l2b_read_state(l2b_ctx* c, int32_t which, int32_t seq, int32_t layer, int32_t pos,
                           float* out, uint64_t n_floats) {
  DevGuard dev_guard;
  if (!c) return L2B_EINVAL;
  if (!out) return fail(c, L2B_EINVAL, "null out");
  if (seq < 0 || seq >= c->Bmax) return fail(c, L2B_EINVAL, "seq %d", seq);
  if (is_group(c)) {
    l2b_ctx* k = is_tp_group(c) ? c->kids[0] : c->kids[seq / c->per_kid];
    return lift(c, k, l2b_read_state(k, which, is_tp_group(c) ? 0 : seq % c->per_kid, layer, pos, out, n_floats));
  }
  if (c->tp_size > 1 && which != L2B_S_LOGITS)
    return fail(c, L2B_ESTATE, "only the logits tap exists in tensor-parallel mode (state is sharded / LL-tagged)");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  const size_t D = c->D;
  const float* src = nullptr;
  size_t n = D;
  switch (which) {
    case L2B_S_X: src = c->x + seq * D; break;
    case L2B_S_Q: src = c->q + seq * D; break;
    case L2B_S_XB: src = c->xb + seq * D; break;
    case L2B_S_HB:
      if (c->last_tc)
        return fail(c, L2B_ESTATE, "hb is not materialised on the tensor-core path (the SwiGLU output exists only "
                                   "as the pre-split operand of the w2 GEMM)");
      src = c->hb + (size_t)seq * c->F; n = c->F; break;
    case L2B_S_LOGITS: src = c->logits + (size_t)seq * c->V; n = c->V; break;
    case L2B_S_KEY_ROW:
    case L2B_S_VALUE_ROW: {
      if (layer < 0 || layer >= c->L || pos < 0 || pos >= c->steps)
        return fail(c, L2B_EINVAL, "layer/pos out of range");
      if (n_floats != D) return fail(c, L2B_EINVAL, "expected %zu floats", D);
      const size_t kv_seq = (size_t)c->H * c->steps * c->hs;
      const float* base = (which == L2B_S_KEY_ROW ? c->kc : c->vc) +
                          ((size_t)layer * c->Bmax + seq) * kv_seq + (size_t)pos * c->hs;
      // head-major [H][steps][hs] -> the reference's row [H*hs]
      CU(c, cudaMemcpy2D(out, c->hs * sizeof(float), base, (size_t)c->steps * c->hs * sizeof(float),
                         c->hs * sizeof(float), c->H, cudaMemcpyDeviceToHost));
      return L2B_OK;
    }
    default: return fail(c, L2B_EINVAL, "which=%d", which);
  }
  if (n_floats != n) return fail(c, L2B_EINVAL, "expected %zu floats, got %llu", n,
                                 (unsigned long long)n_floats);
  CU(c, cudaMemcpy(out, src, n * sizeof(float), cudaMemcpyDeviceToHost));
  return L2B_OK;
}

L2B_API int l2b_debug_timeline(l2b_ctx* c, int64_t* out, uint64_t n) {
  DevGuard dev_guard;
  if (!c || !out) return L2B_EINVAL;
  if (is_group(c)) return lift(c, c->kids[0], l2b_debug_timeline(c->kids[0], out, n));
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  if (c->d_dbg2 && n == 1024 * 12) {   // GEMV launch timeline
    CU(c, cudaMemcpy(out, c->d_dbg2, n * sizeof(long long), cudaMemcpyDeviceToHost));
    return L2B_OK;
  }
  if (!c->d_dbg) return L2B_EINVAL;
  if (n > 256 * 8) n = 256 * 8;
  CU(c, cudaMemcpy(out, c->d_dbg, n * sizeof(long long), cudaMemcpyDeviceToHost));
  return L2B_OK;
}

L2B_API int l2b_reset(l2b_ctx* c) {
  DevGuard dev_guard;
  if (!c) return L2B_EINVAL;
  if (is_group(c)) {
    for (l2b_ctx* k : c->kids) {
      int rc = l2b_reset(k);
      if (rc) return lift(c, k, rc);
    }
    return L2B_OK;
  }
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  const size_t kv = (size_t)c->L * c->Bmax * c->Dl * (size_t)c->steps;
  CU(c, cudaMemsetAsync(c->kc, 0, kv * sizeof(float), c->stream));
  CU(c, cudaMemsetAsync(c->vc, 0, kv * sizeof(float), c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int& v : c->n_run) v = 0;
  return L2B_OK;
}

L2B_API int l2b_set_option(l2b_ctx* c, const char* key, int64_t value) {
  DevGuard dev_guard;
  if (!c || !key) return L2B_EINVAL;
  if (is_group(c)) {
    for (l2b_ctx* k : c->kids) {
      int rc = l2b_set_option(k, key, value);
      if (rc) return lift(c, k, rc);
    }
    return L2B_OK;
  }
  Options& o = c->opt;
  const std::string k(key);
  const int v = (int)value;
  if (k == "graph") o.graph = v != 0;
  else if (k == "pdl") o.pdl = v != 0;
  else if (k == "f64") o.f64 = v != 0;
  else if (k == "threads") {
    if (v != 256 && v != 512) return fail(c, L2B_EINVAL, "threads must be 256 or 512");
    o.threads = v;
  } else if (k == "ctas_per_sm") {
    if (v < 1 || v > 2) return fail(c, L2B_EINVAL, "ctas_per_sm must be 1 or 2");
    o.ctas_per_sm = v;
  } else if (k == "attn_cluster") {
    if (v != 0 && v != 1 && v != 2 && v != 4 && v != 8)
      return fail(c, L2B_EINVAL, "attn_cluster must be 0,1,2,4,8");
    o.attn_cluster = v;
  } else if (k == "evict_first") {
    o.evict_first = v < 0 ? -1 : (v != 0);
  } else if (k == "tc_min_batch") {
    o.tc_min_batch = v < 0 ? 0 : v;
  } else if (k == "l2_prefetch") {
    o.l2_prefetch = v < 0 ? 0 : v;
  } else if (k == "attn_prefetch") {
    o.attn_prefetch = v != 0;
  } else if (k == "soft_sync") {
    o.soft_sync = v != 0;
  } else if (k == "fuse_cluster") {
    if (v != 0 && v != 1 && v != 2 && v != 4 && v != 8) return fail(c, L2B_EINVAL, "fuse_cluster must be 0, 1, 2, 4 or 8");
    o.fuse_cluster = v;
  } else if (k == "fuse_prefetch") {
    o.fuse_prefetch = v < 0 ? 0 : (v > 100 ? 100 : v);
  } else if (k == "fuse_qkv_attn") {
    o.fuse_qkv_attn = v != 0;
  } else if (k == "mega") {
#ifdef L2B_EXPERIMENTS
    o.mega = v < 0 ? 0 : (v > 2 ? 2 : v);
#else
    if (v != 0) return fail(c, L2B_EINVAL, "persistent-kernel experiments are not compiled in (build with -DL2B_EXPERIMENTS)");
#endif
  } else if (k == "stream_stages") {
    o.stream_stages = v < 0 ? 0 : v;
  } else if (k == "stream_chunks") {
    o.stream_chunks = v < 0 ? 0 : v;
  } else if (k == "tp_timeout_ms") {
    if (v < 1) return fail(c, L2B_EINVAL, "tp_timeout_ms must be >= 1");
    o.tp_timeout_ms = v;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    return write_tp_timeout(c);
  } else if (k == "gemv_timeline") {
    if (v && !c->d_dbg2) {
      if (cudaMalloc((void**)&c->d_dbg2, 1024 * 12 * sizeof(long long)) != cudaSuccess) return fail(c, L2B_ENOMEM, "dbg");
    }
    if (c->d_dbg2) cudaMemset(c->d_dbg2, 0, 1024 * 12 * sizeof(long long));
    c->gemv_dbg_arm = v != 0;
    c->gemv_dbg_slot = 0;
  } else if (k == "gemm_timeline") {
    if (v && !c->d_dbg) {
      if (cudaMalloc((void**)&c->d_dbg, 256 * 8 * sizeof(long long)) != cudaSuccess) return fail(c, L2B_ENOMEM, "dbg");
      cudaMemset(c->d_dbg, 0, 256 * 8 * sizeof(long long));
    }
    c->dbg_arm = v != 0;
  } else if (k == "attn_warp") {
    o.attn_warp = v < 0 ? 0 : v;
  } else if (k == "tc_tmem_a") {
    o.tc_tmem_a = v != 0;
  } else if (k == "tc_rewrite_hi") {
    o.tc_rewrite_hi = v != 0;
  } else if (k == "tc_splits") {
    o.tc_splits = v < 0 ? 0 : v;
  } else {
    return fail(c, L2B_EINVAL, "unknown option '%s'", key);
  }
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  drop_graphs(c);
  return L2B_OK;
}

L2B_API int64_t l2b_tp_export(l2b_ctx* c, void* blob, uint64_t cap) {
  DevGuard dev_guard;
  if (!c) return L2B_EINVAL;
  if (is_group(c)) return fail(c, L2B_ESTATE, "a single-process group wires its members itself");
  if (c->tp_size <= 1) return fail(c, L2B_ESTATE, "context was not created with l2b_create_tp");
  if (!blob || cap < sizeof(cudaIpcMemHandle_t)) return fail(c, L2B_EINVAL, "blob needs %zu bytes", sizeof(cudaIpcMemHandle_t));
  CU(c, cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CU(c, cudaIpcGetMemHandle(&h, c->xchg));
  memcpy(blob, &h, sizeof h);
  return (int64_t)sizeof h;
}

L2B_API int l2b_tp_connect(l2b_ctx* c, const void* blobs, uint64_t blob_bytes, int32_t n_ranks) {
  DevGuard dev_guard;
  if (!c) return L2B_EINVAL;
  if (is_group(c)) return fail(c, L2B_ESTATE, "a single-process group wires its members itself");
  if (c->tp_size <= 1) return fail(c, L2B_ESTATE, "context was not created with l2b_create_tp");
  if (!blobs || n_ranks != c->tp_size || blob_bytes < sizeof(cudaIpcMemHandle_t))
    return fail(c, L2B_EINVAL, "expected %d blobs of >= %zu bytes", c->tp_size, sizeof(cudaIpcMemHandle_t));
  CU(c, cudaSetDevice(c->device));
  for (int g = 0; g < c->tp_size; ++g) {
    if (g == c->tp_rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*)blobs + (size_t)g * blob_bytes, sizeof h);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(c, L2B_ECOMM, "cudaIpcOpenMemHandle(rank %d): %s", g, cudaGetErrorString(e));
    }
    c->peer[g] = (unsigned char*)p;
  }
  c->peer_ipc = true;
  c->tp_connected = true;
  drop_graphs(c);
  return L2B_OK;
}

#include "tokenizer.inl"
