// pcu_ctx.cuh -- context: device, stream, reduction workspace, NCCL plumbing.
#pragma once

#include <nccl.h>  // types only; the library is resolved with dlopen at run time

#include <map>
#include <string>
#include <vector>

#include "pcu_common.cuh"
#include "../../include/paropt_b200.h"

struct PendingRed {
  int offset, ns, nx, nm;
};

#define PCU_SHM_CAP 1024  // doubles per publication of the shared-memory all-gather

struct pcu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;

  // fused-kernel reductions
  double *d_partials = nullptr;
  unsigned int *d_counter = nullptr;
  double *d_result = nullptr;   // [PCU_RESULT_CAP]
  double *h_result = nullptr;   // pinned mirror
  double *d_gather = nullptr;   // [world][total] packed partials of all ranks (multi-GPU)
  double *h_gather = nullptr;   // pinned mirror, combined on the host in rank order
  // zero-copy results (RedBuf::hres / hflag): page-locked host memory mapped into the
  // device; [0, CAP + MAX_RED) doubles of results, then the flag word
  double *h_zc = nullptr, *d_zc = nullptr;
  unsigned long long *h_zflag = nullptr, *d_zflag = nullptr;
  unsigned long long zc_seq = 0;       // sequence number of the last reduction launched
  bool zc_on = false;
  // single-node all-gather of the ranks' packed partials through POSIX shared memory
  // (the values are consumed by the hosts: PCIe write + shared memory is the shortest
  // path from N GPUs to N host threads); NCCL all-gather when unavailable
  struct ShmRank {
    volatile unsigned long long seq;
    char pad[56];
    double data[2][PCU_SHM_CAP];
  };
  void *shm_base = nullptr;
  size_t shm_bytes = 0;
  unsigned long long shm_pub = 0;      // publications made so far
  char shm_name[64] = {0};
  int wait_flag(unsigned long long seq);
  int shm_setup(const unsigned char id128[128]);
  int shm_allgather(int total, const double *src = nullptr, double *dst = nullptr);
  std::vector<double> big_gather;      // [world][n] of big_fetch's host-side all-reduce
  int result_used = 0;
  bool red_overflow = false;   // redbuf() ran out of slots: the next fetch() fails
  std::vector<PendingRed> pending;

  // large sum-reductions (mdot, Gram triangle)
  double *d_big = nullptr;
  double *h_big = nullptr;
  size_t big_cap = 0;
  double *d_big_partials = nullptr;
  size_t big_partials_cap = 0;

  int grid = 148 * 4;
  int num_sms = 148;
  int max_blocks_per_sm = 16;  // occupancy cap of the streaming kernels
  int prefetch = -1;           // -1: same-iteration L2 prefetch; k > 0: k iterations ahead; 0 off
  int no_shm_big = 1;          // 0 (PCU_SHM_BIG=1): big_fetch adds the ranks' partials on the hosts
  int no_reverse = 0;          // PCU_NO_REVERSE: every staged pass walks its tiles upwards
  int no_tma_tile = 0;         // PCU_NO_TMA_TILE: keep the SRC functors on the register-fed kernel
  int tma_groups = 0;          // PCU_TMA_GROUPS: cap on the consumer groups of tma_tile_kernel
  int tma_npw = 0;             // PCU_TMA_NPW: producer warps of tma_tile_kernel (default 2)
  int tma_min_tiles = 0;       // staged launch only from this many tiles (default 8 per SM)
  int tma_grid = 0;            // CTAs of the staged launch (default: one per SM)
  int tma_max_rows = 0;        // debugging: staged launch only for tiles up to this many rows
  int managed_vectors = 0;     // pcu_vec_create allocates unified memory (pcu_vec_host_ptr)
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  // optional per-kernel device timing (CUDA events on the launching stream)
  struct ProfPending {
    const char *name;
    cudaEvent_t e0, e1;
  };
  struct ProfTotal {
    double ms = 0.0;
    long count = 0;
  };
  bool profiling = false;
  std::vector<ProfPending> prof_pending;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t> > prof_pool;
  std::map<std::string, ProfTotal> prof_totals;
  void prof_begin(const char *name);
  void prof_end();
  void prof_collect();  // call after a stream synchronisation

  // Deferred results: defer() marks every slot reserved so far; the next fetch() -- the
  // caller's own or one made by a problem callback in between -- brings them to the
  // host in the same copy and keeps them aside for take_deferred(), handing its caller
  // only the slots reserved after the mark.  One synchronisation serves both.
  int deferred_n = 0;
  bool deferred_ready = false;
  std::vector<double> deferred_vals;
  void defer() { deferred_n = result_used; deferred_ready = false; }
  int take_deferred(double *out, int n);

  // the device chain of the KKT solve keeps its coefficient tables in a constant bank
  // (one per device and process): only while this is the device's only context
  bool chain_ok() const;

  RedBuf redbuf(int ns, int nx, int nm);       // reserves a result slot
  int fetch(double *out);                      // all pending slots -> host, sync
  int big_reserve(size_t nresult, size_t npartials);
  int big_fetch(size_t n, double *out);        // allreduce(sum) + D2H + sync
};

#define PCU_RESULT_CAP 512

// NCCL entry points resolved at run time (torch's bundled libnccl.so.2 when the
// process already loaded it, the system one otherwise).
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t,
                            ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t,
                            ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi &nccl_api();

struct pcu_vec {
  pcu_ctx *ctx = nullptr;
  int n = 0;
  double *d = nullptr;
  bool owns = true;
  bool managed = false;       // unified memory (context parameter "managed_vectors")
  bool host_touched = false;  // handed to the host through pcu_vec_host_ptr since the last use
};

// Every C-ABI entry point that reads or writes a caller's vector on the device
// passes it through here first: a managed vector the host has touched is moved
// back to the device on the context's stream (pcu_vec_host_ptr, paropt_b200.h).
static inline void pcu_vec_ready(pcu_vec *v) {
  if (v && v->host_touched) {
    v->host_touched = false;
    if (v->n > 0)
      cudaMemPrefetchAsync(v->d, ((size_t)v->n + 2) * sizeof(double), v->ctx->device,
                           v->ctx->stream);
  }
}

template <class F>
const char *pcu_kernel_name() {
  return __PRETTY_FUNCTION__;  // "... [with F = ResF]"; trimmed when reported
}

// launch helper: grid sized to the work, capped at the persistent grid
static inline int pcu_grid_for(const pcu_ctx *ctx, long long n) {
  long long need = (n / 2 + PCU_THREADS - 1) / PCU_THREADS;
  if (need < 1) need = 1;
  if (need > ctx->grid) need = ctx->grid;
  return (int)need;
}

// Host-side visit of a functor's staged streams: which slots are live, alignment.
struct TmaHostCheck {
  int nfix = 0;
  int copies = 0, ncols = 0;
  bool aligned = true;
  bool fixed[24] = {false};
  __host__ __device__ void n(int slot, const double *ptr) {
    if (!ptr) return;
    copies++;
    if (((uintptr_t)ptr) & 15) aligned = false;
    if (slot < nfix) fixed[slot] = true;
    else if (slot - nfix + 1 > ncols) ncols = slot - nfix + 1;
  }
  __host__ __device__ void w(int, const double *ptr) {
    if (!ptr) return;
    copies++;
    if (((uintptr_t)ptr) & 15) aligned = false;
  }
};

#define PCU_MAX_DEVICES 64
#define PCU_TMA_SMEM_BUDGET (222 * 1024)  // + <= 5 KB static: the 227 KB a CTA may hold

// Bulk-copy staged launch of a SRC functor (tma_tile_kernel); returns -1 when
// the launch does not qualify (small n, generic weighting pattern, too many
// streams for the shared-memory ring) and the register-fed kernel must run.
template <class F>
int pcu_launch_tile_tma(pcu_ctx *ctx, const F &f, long long n, const WDesc &w,
                        RedBuf rb) {
  constexpr int ROWS = F::TROWS, WPT = PCU_TMA_WPT;
  static_assert(F::NFIX <= 24 && ROWS % (64 * WPT) == 0, "staged tile shape");
  if (ctx->no_tma_tile || w.mode == 2) return -1;
  if (ctx->tma_max_rows > 0 && ROWS > ctx->tma_max_rows) return -1;
  if (w.mode == 1 && (w.nw < 2 || w.nw > 64 || (w.nw & (w.nw - 1)) || w.wstart != 0 ||
                      w.wstride != w.nw))
    return -1;
  TmaPlan plan;
  plan.ntiles = n / ROWS;
  if (plan.ntiles < (ctx->tma_min_tiles > 0 ? (long long)ctx->tma_min_tiles
                                            : (long long)ctx->num_sms * 8))
    return -1;
  TmaHostCheck chk;
  chk.nfix = F::NFIX;
  f.tstreams(chk);
  int nslots = 0;  // live fixed slots, compacted, then the columns
  plan.nmap[0] = plan.nmap[1] = plan.nmap[2] = 0ull;
  for (int i = 0; i < 24; i++) plan.noff[i] = 0u;
  for (int i = 0; i < F::NFIX; i++) {
    plan.nmap[i >> 3] |= (unsigned long long)nslots << ((i & 7) * 8);
    plan.noff[i] = (unsigned)nslots * (unsigned)(ROWS * 8);
    if (chk.fixed[i]) nslots++;
  }
  plan.col_base = nslots;
  plan.reverse = (F::REVERSE && !ctx->no_reverse) ? 1 : 0;
  nslots += chk.ncols;
  int npw = ctx->tma_npw > 0 ? ctx->tma_npw : 4;
  if (npw > PCU_TMA_NPW) npw = PCU_TMA_NPW;
  plan.npw = npw;
  if (!chk.aligned || chk.copies > 64 * npw) return -1;
  const int con_per_tile = w.mode == 1 ? ROWS / w.nw : 0;
  plan.wpitch = (con_per_tile * 8 + 15) / 16 * 16;
  plan.woff = nslots * ROWS * 8;
  plan.stage_bytes = (plan.woff + (w.mode == 1 ? F::NWSLOTS * plan.wpitch : 0) + 127) / 128 * 128;
  const int fit = PCU_TMA_SMEM_BUDGET / plan.stage_bytes;
  int groups = (PCU_TMA_MAXWARPS - PCU_TMA_NPW) / WPT;
  if (ctx->tma_groups > 0 && ctx->tma_groups < groups) groups = ctx->tma_groups;
  if (fit < groups) groups = fit;
  if (groups < 2) return -1;
  int depth = fit / groups;
  if (depth > PCU_TMA_MAXSTAGES / groups) depth = PCU_TMA_MAXSTAGES / groups;
  plan.groups = groups;
  plan.nstages = groups * depth;
  const long long ncon_elems = w.mode == 1 ? (long long)w.nwcon * w.nw : 0;
  plan.tiles_con = ncon_elems / ROWS;
  if (plan.tiles_con > plan.ntiles) plan.tiles_con = plan.ntiles;
  plan.tile_skip = (ncon_elems % ROWS != 0 && plan.tiles_con < plan.ntiles) ? plan.tiles_con : -1;
  // function attributes are per device: one flag per ordinal (several contexts /
  // devices may live in one process)
  static bool attr_set[PCU_MAX_DEVICES] = {false};
  const int dev = ctx->device >= 0 && ctx->device < PCU_MAX_DEVICES ? ctx->device : 0;
  if (!attr_set[dev] || ctx->device >= PCU_MAX_DEVICES) {
    PCU_CUDA_OK(cudaFuncSetAttribute(tma_tile_kernel<F, ROWS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     PCU_TMA_SMEM_BUDGET));
    attr_set[dev] = true;
  }
  const int threads = PCU_TMA_MAXWARPS * 32;
  const size_t smem = (size_t)plan.nstages * plan.stage_bytes;
  ctx->prof_begin(pcu_kernel_name<F>());
  const int grid = ctx->tma_grid > 0 && ctx->tma_grid < ctx->num_sms ? ctx->tma_grid : ctx->num_sms;
  tma_tile_kernel<F, ROWS><<<grid, threads, smem, ctx->stream>>>(f, n, w, rb, plan);
  ctx->prof_end();
  ctx->launches++;
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

template <class F>
int pcu_launch_tile(pcu_ctx *ctx, const F &f, long long n, const WDesc &w,
                    RedBuf rb) {
  if constexpr (F::SRC) {
    const int r = pcu_launch_tile_tma(ctx, f, n, w, rb);
    if (r >= 0) return r;
  }
  // persistent grid: exactly as many blocks as can be co-resident
  static int bps_dev[PCU_MAX_DEVICES];
  static bool bps_init = false;
  if (!bps_init) {
    for (int i = 0; i < PCU_MAX_DEVICES; i++) bps_dev[i] = -1;
    bps_init = true;
  }
  const int dev_ = ctx->device >= 0 && ctx->device < PCU_MAX_DEVICES ? ctx->device : 0;
  if (ctx->device >= PCU_MAX_DEVICES) bps_dev[dev_] = -1;  // beyond the table: always query
  int &blocks_per_sm = bps_dev[dev_];
  if (blocks_per_sm < 0) {
    int v = 0;
    if (F::SMEM > 0)
      cudaFuncSetAttribute(tile_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           F::SMEM);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, tile_kernel<F>,
                                                      PCU_TILE_THREADS, F::SMEM) != cudaSuccess ||
        v < 1)
      v = 1;
    blocks_per_sm = v > 16 ? 16 : v;
  }
  long long need = (n / 2 + PCU_TILE_THREADS - 1) / PCU_TILE_THREADS;
  if (need < 1) need = 1;
  int bps = blocks_per_sm < ctx->max_blocks_per_sm ? blocks_per_sm : ctx->max_blocks_per_sm;
  long long cap = (long long)ctx->num_sms * bps;
  if (cap > PCU_MAX_BLOCKS) cap = PCU_MAX_BLOCKS;
  const int grid = (int)(need < cap ? need : cap);
  ctx->prof_begin(pcu_kernel_name<F>());
  rb.prefetch = ctx->prefetch;
  tile_kernel<F><<<grid, PCU_TILE_THREADS, F::SMEM, ctx->stream>>>(f, n, w, rb);
  ctx->prof_end();
  ctx->launches++;
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}
