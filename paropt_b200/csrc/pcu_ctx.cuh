// pcu_ctx.cuh -- context: device, stream, reduction workspace, NCCL plumbing.
#pragma once

#include <nccl.h>  // types only; the library is resolved with dlopen at run time

#include <vector>

#include "pcu_common.cuh"
#include "../../include/paropt_b200.h"

struct PendingRed {
  int offset, ns, nx, nm;
};

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
  double *d_gather = nullptr;   // [world][PCU_RESULT_CAP] (multi-GPU)
  int result_used = 0;
  std::vector<PendingRed> pending;

  // large sum-reductions (mdot, Gram triangle)
  double *d_big = nullptr;
  double *h_big = nullptr;
  size_t big_cap = 0;
  double *d_big_partials = nullptr;
  size_t big_partials_cap = 0;

  int grid = 148 * 4;
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

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
};

// launch helper: grid sized to the work, capped at the persistent grid
static inline int pcu_grid_for(const pcu_ctx *ctx, long long n) {
  long long need = (n / 2 + PCU_THREADS - 1) / PCU_THREADS;
  if (need < 1) need = 1;
  if (need > ctx->grid) need = ctx->grid;
  return (int)need;
}
