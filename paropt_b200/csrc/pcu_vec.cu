// pcu_vec.cu -- context implementation and the ParOptVec kernels
// (reference: src/ParOptVec.cpp:15-217; one fused, deterministic kernel per
// BLAS-1 call + MPI_Allreduce pair of the reference).
#include <dlfcn.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>

#include "pcu_ctx.cuh"

// ------------------------------------------------------------------ NCCL api
NcclApi &nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
      api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
      api.GetErrorString =
          (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce &&
               api.AllGather;
    }
  }
  return api;
}

#define PCU_NCCL_OK(call)                                                     \
  do {                                                                        \
    ncclResult_t r__ = (call);                                                \
    if (r__ != ncclSuccess) {                                                 \
      fprintf(stderr, "paropt_b200: NCCL error %d at %s:%d\n", (int)r__,      \
              __FILE__, __LINE__);                                            \
      return 1;                                                               \
    }                                                                         \
  } while (0)

static int g_ctx_on_device[PCU_MAX_DEVICES] = {0};
bool pcu_ctx::chain_ok() const {
  return device >= 0 && device < PCU_MAX_DEVICES && g_ctx_on_device[device] == 1;
}

RedBuf pcu_ctx::redbuf(int ns, int nx, int nm) {
  RedBuf rb = {};
  rb.prefetch = 0;
  rb.partials = d_partials;
  rb.counter = d_counter;
  const int nr = ns + nx + nm;
  if (result_used + nr > PCU_RESULT_CAP || pending.size() >= 32) {
    // a caller enqueued reductions without ever fetching them: fail the next fetch
    // (pending results are kept) and park this kernel's output in the spare slots
    if (!red_overflow)
      fprintf(stderr, "paropt_b200: reduction slots exhausted (%d pending): the next fetch fails\n",
              (int)pending.size());
    red_overflow = true;
    rb.result = d_result + PCU_RESULT_CAP;
    return rb;
  }
  rb.result = d_result + result_used;
  if (zc_on) {
    rb.hres = d_zc + result_used;
    rb.hflag = d_zflag;
    rb.seq = ++zc_seq;
  }
  pending.push_back({result_used, ns, nx, nm});
  result_used += nr;
  return rb;
}

// Host side of the zero-copy hand-over: poll the flag the last block of the most
// recent reduction kernel publishes (kernels of one stream complete in order).
int pcu_ctx::wait_flag(unsigned long long seq) {
  volatile unsigned long long *flag = h_zflag;
  const auto t0 = std::chrono::steady_clock::now();
  unsigned long spins = 0;
  while (*flag < seq) {
    if ((++spins & 0xfff) == 0) {
      const cudaError_t q = cudaStreamQuery(stream);
      if (q != cudaSuccess && q != cudaErrorNotReady) {
        fprintf(stderr, "paropt_b200: CUDA error while waiting for a reduction: %s\n",
                cudaGetErrorString(q));
        return 1;
      }
      if (q == cudaSuccess && *flag < seq) {
        // the stream drained without the flag: the launch itself failed
        if (*flag < seq) {
          fprintf(stderr, "paropt_b200: reduction result never arrived\n");
          return 1;
        }
      }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
        fprintf(stderr, "paropt_b200: timed out waiting for a reduction\n");
        return 1;
      }
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return 0;
}

// ----- shared-memory all-gather between the ranks of one node
static unsigned long long hash128(const unsigned char id[128]) {
  unsigned long long h = 1469598103934665603ull;
  for (int i = 0; i < 128; i++) h = (h ^ id[i]) * 1099511628211ull;
  return h;
}

int pcu_ctx::shm_setup(const unsigned char id128[128]) {
  snprintf(shm_name, sizeof(shm_name), "/pcu_%016llx", hash128(id128));
  shm_bytes = 4096 + sizeof(ShmRank) * (size_t)world;
  int fd = -1;
  if (rank == 0) {
    shm_unlink(shm_name);
    fd = shm_open(shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd >= 0 && ftruncate(fd, (off_t)shm_bytes) != 0) {
      close(fd);
      fd = -1;
    }
  } else {
    for (int tries = 0; tries < 20000 && fd < 0; tries++) {  // rank 0 creates it first
      fd = shm_open(shm_name, O_RDWR, 0600);
      if (fd >= 0) {
        struct stat st;
        if (fstat(fd, &st) != 0 || (size_t)st.st_size < shm_bytes) {
          close(fd);
          fd = -1;
        }
      }
      if (fd < 0) usleep(500);
    }
  }
  if (fd < 0) return 1;
  void *p = mmap(nullptr, shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) return 1;
  shm_base = p;
  // attach count: rank 0 removes the name once everybody is in
  volatile int *attached = reinterpret_cast<volatile int *>(p);
  __sync_fetch_and_add(attached, 1);
  if (rank == 0) {
    for (long tries = 0; *attached < world && tries < 40000; tries++) usleep(500);
    shm_unlink(shm_name);
    if (*attached < world) return 1;
  }
  return 0;
}

// src / dst default to the zero-copy result buffer and h_gather
int pcu_ctx::shm_allgather(int total, const double *src, double *dst) {
  if (!src) src = h_zc;
  if (!dst) dst = h_gather;
  ShmRank *ranks = reinterpret_cast<ShmRank *>(reinterpret_cast<char *>(shm_base) + 4096);
  const unsigned long long n = ++shm_pub;
  ShmRank &mine = ranks[rank];
  memcpy(mine.data[n & 1], src, sizeof(double) * (size_t)total);
  std::atomic_thread_fence(std::memory_order_release);
  mine.seq = n;
  const auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < world; r++) {
    unsigned long spins = 0;
    while (ranks[r].seq < n) {
      if ((++spins & 0xffff) == 0 &&
          std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
        fprintf(stderr, "paropt_b200: rank %d timed out waiting for rank %d\n", rank, r);
        return 1;
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    memcpy(dst + (size_t)r * total, ranks[r].data[n & 1], sizeof(double) * (size_t)total);
  }
  return 0;
}

int pcu_ctx::fetch(double *out) {
  if (red_overflow) {
    red_overflow = false;
    result_used = 0;
    deferred_n = 0;
    pending.clear();
    cudaStreamSynchronize(stream);
    return 1;
  }
  const int total = result_used;
  const bool zc = zc_on && total > 0 && total <= 512;
  const bool via_shm = zc && world > 1 && shm_base != nullptr;
  if (zc && (world == 1 || via_shm)) {
    // zero-copy: the results are already on their way into h_zc
    if (wait_flag(zc_seq)) return 1;
    if (world == 1) memcpy(h_result, h_zc, sizeof(double) * (size_t)total);
    else if (shm_allgather(total)) return 1;
  } else if (total > 0) {
    if (world > 1) {
      // one all-gather of the `total` packed partials; the fixed-rank-order combine
      // runs on the host of every rank (identical data, identical order: identical
      // results) -- no combine kernel between the collective and the copy
      NcclApi &api = nccl_api();
      PCU_NCCL_OK(api.AllGather(d_result, d_gather, (size_t)total, ncclFloat64, comm, stream));
      PCU_CUDA_OK(cudaMemcpyAsync(h_gather, d_gather, (size_t)world * total * sizeof(double),
                                  cudaMemcpyDeviceToHost, stream));
    } else {
      PCU_CUDA_OK(cudaMemcpyAsync(h_result, d_result, total * sizeof(double),
                                  cudaMemcpyDeviceToHost, stream));
    }
  }
  if (!(zc && (world == 1 || via_shm))) PCU_CUDA_OK(cudaStreamSynchronize(stream));
  if (world > 1 && total > 0) {
    for (const PendingRed &pr : pending) {
      const int nr = pr.ns + pr.nx + pr.nm;
      for (int i = 0; i < nr; i++) {
        const int idx = pr.offset + i;
        double v = h_gather[idx];
        for (int r = 1; r < world; r++) {
          const double p = h_gather[(size_t)r * total + idx];
          if (i < pr.ns) v += p;
          else if (i < pr.ns + pr.nx) v = fmax(v, p);
          else v = fmin(v, p);
        }
        h_result[idx] = v;
      }
    }
  }
  int skip = 0;
  if (deferred_n > 0) {  // results of an earlier launch ride along (see defer())
    skip = deferred_n <= total ? deferred_n : total;
    deferred_vals.assign(h_result, h_result + skip);
    deferred_ready = true;
    deferred_n = 0;
  }
  if (out && total > skip) memcpy(out, h_result + skip, (total - skip) * sizeof(double));
  result_used = 0;
  pending.clear();
  return 0;
}

int pcu_ctx::take_deferred(double *out, int n) {
  if (!deferred_ready) {
    // nobody fetched since defer(): the deferred slots are still on the device
    if (deferred_n <= 0 || fetch(nullptr)) return 1;
  }
  deferred_ready = false;
  if ((int)deferred_vals.size() < n) return 1;
  memcpy(out, deferred_vals.data() + (deferred_vals.size() - n), n * sizeof(double));
  return 0;
}

void pcu_ctx::prof_begin(const char *name) {
  if (!profiling) return;
  if (prof_pool.empty()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    prof_pool.push_back({a, b});
  }
  ProfPending p;
  p.name = name;
  p.e0 = prof_pool.back().first;
  p.e1 = prof_pool.back().second;
  prof_pool.pop_back();
  cudaEventRecord(p.e0, stream);
  prof_pending.push_back(p);
}

void pcu_ctx::prof_end() {
  if (!profiling || prof_pending.empty()) return;
  cudaEventRecord(prof_pending.back().e1, stream);
}

void pcu_ctx::prof_collect() {
  for (auto &p : prof_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.e1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
      std::string name(p.name);
      size_t pos = name.find("F = ");
      if (pos != std::string::npos) {
        name = name.substr(pos + 4);
        size_t end = name.find_first_of("];");
        if (end != std::string::npos) name = name.substr(0, end);
      }
      ProfTotal &t = prof_totals[name];
      t.ms += ms;
      t.count += 1;
    }
    prof_pool.push_back({p.e0, p.e1});
  }
  prof_pending.clear();
}

int pcu_ctx::big_reserve(size_t nresult, size_t npartials) {
  if (nresult > big_cap) {
    if (d_big) cudaFree(d_big);
    if (h_big) cudaFreeHost(h_big);
    big_cap = nresult * 2;
    PCU_CUDA_OK(cudaMalloc(&d_big, big_cap * sizeof(double)));
    PCU_CUDA_OK(cudaMallocHost(&h_big, big_cap * sizeof(double)));
  }
  if (npartials > big_partials_cap) {
    if (d_big_partials) cudaFree(d_big_partials);
    big_partials_cap = npartials * 2;
    PCU_CUDA_OK(cudaMalloc(&d_big_partials, big_partials_cap * sizeof(double)));
  }
  return 0;
}

int pcu_ctx::big_fetch(size_t n, double *out) {
  if (world > 1 && n > 0 && n <= PCU_SHM_CAP && shm_base != nullptr && !no_shm_big) {
    // ranks of one node: every rank copies its partial to the host, the partials are
    // exchanged through the shared-memory segment and added in rank order on every host
    // (identical data, identical order: identical results) -- the result is consumed by
    // the host anyway, and the in-stream all-reduce of a few hundred doubles costs
    // 20-30 us of latency at 8 GPUs.  d_big keeps the LOCAL partial.
    PCU_CUDA_OK(cudaMemcpyAsync(h_big, d_big, n * sizeof(double), cudaMemcpyDeviceToHost,
                                stream));
    PCU_CUDA_OK(cudaStreamSynchronize(stream));
    big_gather.resize((size_t)world * n);
    if (shm_allgather((int)n, h_big, big_gather.data())) return 1;
    for (size_t i = 0; i < n; i++) {
      double v = big_gather[i];
      for (int r = 1; r < world; r++) v += big_gather[(size_t)r * n + i];
      h_big[i] = v;
    }
    if (out) memcpy(out, h_big, n * sizeof(double));
    return 0;
  }
  if (world > 1 && n > 0) {
    PCU_NCCL_OK(nccl_api().AllReduce(d_big, d_big, n, ncclFloat64, ncclSum, comm,
                                     stream));
  }
  if (n > 0) {
    PCU_CUDA_OK(cudaMemcpyAsync(h_big, d_big, n * sizeof(double),
                                cudaMemcpyDeviceToHost, stream));
  }
  PCU_CUDA_OK(cudaStreamSynchronize(stream));
  if (out && n > 0) memcpy(out, h_big, n * sizeof(double));
  return 0;
}

// ------------------------------------------------------------- C ABI: context
extern "C" {

int pcu_ctx_allreduce_sum(pcu_ctx *ctx, double *vals, int n) {
  if (!ctx || n < 0) return 1;
  if (ctx->world <= 1 || n == 0) return 0;
  if (ctx->big_reserve((size_t)n, 0)) return 1;
  memcpy(ctx->h_big, vals, sizeof(double) * (size_t)n);
  PCU_CUDA_OK(cudaMemcpyAsync(ctx->d_big, ctx->h_big, sizeof(double) * (size_t)n,
                              cudaMemcpyHostToDevice, ctx->stream));
  return ctx->big_fetch((size_t)n, vals);
}

const char *pcu_version(void) { return "paropt_b200 0.1 (sm_100a)"; }

pcu_ctx *pcu_ctx_create(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr,
            "paropt_b200: no CUDA device available (there is no CPU path)\n");
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    fprintf(stderr, "paropt_b200: cannot select CUDA device %d\n", device);
    return nullptr;
  }
  pcu_ctx *ctx = new pcu_ctx;
  ctx->device = device;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->num_sms = prop.multiProcessorCount;
  if (const char *e = getenv("PCU_MAX_BLOCKS_PER_SM")) ctx->max_blocks_per_sm = atoi(e);
  if (const char *e = getenv("PCU_PREFETCH")) ctx->prefetch = atoi(e);
  if (getenv("PCU_NO_TMA_TILE")) ctx->no_tma_tile = 1;
  if (getenv("PCU_NO_REVERSE")) ctx->no_reverse = 1;
  // measured at 2 GPUs: 6.81 ms / iteration against 6.77 with the in-stream NCCL
  // all-reduce -- no gain, so the host-side variant is opt-in
  ctx->no_shm_big = getenv("PCU_SHM_BIG") ? 0 : 1;
  if (const char *e = getenv("PCU_TMA_GROUPS")) ctx->tma_groups = atoi(e);
  if (const char *e = getenv("PCU_TMA_NPW")) ctx->tma_npw = atoi(e);
  if (const char *e = getenv("PCU_TMA_MIN_TILES")) ctx->tma_min_tiles = atoi(e);
  if (const char *e = getenv("PCU_TMA_GRID")) ctx->tma_grid = atoi(e);
  ctx->grid = prop.multiProcessorCount * 4;
  if (ctx->grid > PCU_MAX_BLOCKS) ctx->grid = PCU_MAX_BLOCKS;
  bool ok = true;
  ok &= cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok &= cudaMalloc(&ctx->d_partials, sizeof(double) * PCU_MAX_BLOCKS * PCU_MAX_RED) == cudaSuccess;
  ok &= cudaMalloc(&ctx->d_counter, 64) == cudaSuccess;
  ok &= cudaMalloc(&ctx->d_result, sizeof(double) * (PCU_RESULT_CAP + PCU_MAX_RED)) == cudaSuccess;
  ok &= cudaMallocHost(&ctx->h_result, sizeof(double) * PCU_RESULT_CAP) == cudaSuccess;
  if (!getenv("PCU_NO_ZEROCOPY")) {
    void *hp = nullptr, *dp = nullptr;
    const size_t zbytes = sizeof(double) * (PCU_RESULT_CAP + PCU_MAX_RED) + 128;
    if (cudaHostAlloc(&hp, zbytes, cudaHostAllocMapped) == cudaSuccess &&
        cudaHostGetDevicePointer(&dp, hp, 0) == cudaSuccess) {
      memset(hp, 0, zbytes);
      ctx->h_zc = (double *)hp;
      ctx->d_zc = (double *)dp;
      ctx->h_zflag = (unsigned long long *)((char *)hp + zbytes - 64);
      ctx->d_zflag = (unsigned long long *)((char *)dp + zbytes - 64);
      ctx->zc_on = true;
    }
  }
  ok &= cudaMemset(ctx->d_counter, 0, 64) == cudaSuccess;
  ok &= cudaMemset(ctx->d_result, 0, sizeof(double) * (PCU_RESULT_CAP + PCU_MAX_RED)) == cudaSuccess;
  ok &= cudaEventCreate(&ctx->ev0) == cudaSuccess;
  ok &= cudaEventCreate(&ctx->ev1) == cudaSuccess;
  ok &= cudaDeviceSynchronize() == cudaSuccess;
  if (!ok) {
    fprintf(stderr, "paropt_b200: context allocation failed: %s\n",
            cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return nullptr;
  }
  if (device >= 0 && device < PCU_MAX_DEVICES) g_ctx_on_device[device]++;
  return ctx;
}

void pcu_ctx_destroy(pcu_ctx *ctx) {
  if (!ctx) return;
  if (ctx->device >= 0 && ctx->device < PCU_MAX_DEVICES) g_ctx_on_device[ctx->device]--;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm && nccl_api().CommDestroy) nccl_api().CommDestroy(ctx->comm);
  cudaFree(ctx->d_partials);
  cudaFree(ctx->d_counter);
  cudaFree(ctx->d_result);
  cudaFreeHost(ctx->h_result);
  if (ctx->h_zc) cudaFreeHost(ctx->h_zc);
  if (ctx->shm_base) munmap(ctx->shm_base, ctx->shm_bytes);
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
  if (ctx->d_big) cudaFree(ctx->d_big);
  if (ctx->h_big) cudaFreeHost(ctx->h_big);
  if (ctx->d_big_partials) cudaFree(ctx->d_big_partials);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int pcu_nccl_unique_id(unsigned char id128[128]) {
  NcclApi &api = nccl_api();
  if (!api.ok) {
    fprintf(stderr, "paropt_b200: libnccl.so.2 not found\n");
    return 1;
  }
  ncclUniqueId id;
  PCU_NCCL_OK(api.GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return 0;
}

int pcu_ctx_init_comm(pcu_ctx *ctx, const unsigned char id128[128], int rank,
                      int world_size) {
  if (world_size <= 1) {
    ctx->rank = 0;
    ctx->world = 1;
    return 0;
  }
  NcclApi &api = nccl_api();
  if (!api.ok) {
    fprintf(stderr, "paropt_b200: libnccl.so.2 not found\n");
    return 1;
  }
  PCU_CUDA_OK(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  PCU_NCCL_OK(api.CommInitRank(&ctx->comm, world_size, id, rank));
  ctx->rank = rank;
  ctx->world = world_size;
  PCU_CUDA_OK(cudaMalloc(&ctx->d_gather,
                         sizeof(double) * PCU_RESULT_CAP * world_size));
  PCU_CUDA_OK(cudaMallocHost(&ctx->h_gather,
                             sizeof(double) * PCU_RESULT_CAP * world_size));
  // host-consumed reductions: shared-memory all-gather between the ranks of this node
  // (every rank decides alike: the verdict is all-reduced over NCCL)
  int have = 0;
  if (ctx->zc_on && !getenv("PCU_NO_SHM")) have = ctx->shm_setup(id128) == 0 ? 1 : 0;
  int *d_have = nullptr;
  PCU_CUDA_OK(cudaMalloc(&d_have, sizeof(int)));
  PCU_CUDA_OK(cudaMemcpy(d_have, &have, sizeof(int), cudaMemcpyHostToDevice));
  PCU_NCCL_OK(api.AllReduce(d_have, d_have, 1, ncclInt32, ncclMin, ctx->comm, ctx->stream));
  PCU_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  PCU_CUDA_OK(cudaMemcpy(&have, d_have, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(d_have);
  if (!have && ctx->shm_base) {
    munmap(ctx->shm_base, ctx->shm_bytes);
    ctx->shm_base = nullptr;
  }
  return 0;
}

int pcu_ctx_rank(pcu_ctx *ctx) { return ctx->rank; }
int pcu_ctx_size(pcu_ctx *ctx) { return ctx->world; }
int pcu_ctx_sync(pcu_ctx *ctx) {
  PCU_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
void *pcu_ctx_stream(pcu_ctx *ctx) { return (void *)ctx->stream; }
int64_t pcu_ctx_kernel_launches(pcu_ctx *ctx) { return ctx->launches; }

int pcu_ctx_set_param(pcu_ctx *ctx, const char *name, int value) {
  if (!ctx || !name) return 1;
  const std::string k(name);
  if (k == "prefetch") ctx->prefetch = value;
  else if (k == "max_blocks_per_sm") ctx->max_blocks_per_sm = value;
  else if (k == "no_tma_tile") ctx->no_tma_tile = value;
  else if (k == "tma_groups") ctx->tma_groups = value;
  else if (k == "tma_npw") ctx->tma_npw = value;
  else if (k == "tma_min_tiles") ctx->tma_min_tiles = value;
  else if (k == "tma_grid") ctx->tma_grid = value;
  else if (k == "tma_max_rows") ctx->tma_max_rows = value;
  else if (k == "managed_vectors") ctx->managed_vectors = value;
  else return 1;
  return 0;
}
int pcu_ctx_profile(pcu_ctx *ctx, int enable) {
  cudaStreamSynchronize(ctx->stream);
  ctx->prof_collect();
  if (enable == 2) ctx->prof_totals.clear();
  ctx->profiling = enable != 0;
  return 0;
}
int pcu_ctx_profile_count(pcu_ctx *ctx) {
  cudaStreamSynchronize(ctx->stream);
  ctx->prof_collect();
  return (int)ctx->prof_totals.size();
}
int pcu_ctx_profile_get(pcu_ctx *ctx, int index, char *name, int name_len,
                        double *ms, int64_t *count) {
  int i = 0;
  for (auto &kv : ctx->prof_totals) {
    if (i++ == index) {
      snprintf(name, name_len, "%s", kv.first.c_str());
      *ms = kv.second.ms;
      *count = kv.second.count;
      return 0;
    }
  }
  return 1;
}
int pcu_ctx_timer_start(pcu_ctx *ctx) {
  PCU_CUDA_OK(cudaEventRecord(ctx->ev0, ctx->stream));
  return 0;
}
int pcu_ctx_timer_stop(pcu_ctx *ctx, double *ms) {
  PCU_CUDA_OK(cudaEventRecord(ctx->ev1, ctx->stream));
  PCU_CUDA_OK(cudaEventSynchronize(ctx->ev1));
  float f = 0.f;
  PCU_CUDA_OK(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
  *ms = (double)f;
  return 0;
}

}  // extern "C"

// --------------------------------------------------------------- vec kernels
struct NoElem {};
struct NoCon {
  static constexpr int ND = 0;
  double d[1];
  __device__ __forceinline__ void zero() {}
};

// y = alpha  |  y *= alpha  |  y += alpha * x      (ParOptVec.cpp:32,177,194)
struct VecOpF : NoStreams {
  static constexpr int NS = 0, NX = 0, NM = 0, NB = 0;
  typedef Acc<NS, NX, NM> AccT;
  typedef NoElem Elem;
  typedef NoCon Con;
  int op;  // 0 set, 1 scale, 2 axpy
  double alpha;
  double *y;
  const double *x;
  template <int W>
  __device__ __forceinline__ void A(long long, const double (&)[W], Elem (&)[W],
                                    double (&)[W][1], AccT *acc) const {}
  __device__ __forceinline__ void B(long long, const double (&)[1], Con &,
                                    AccT &) const {}
  template <int W>
  __device__ __forceinline__ void C(long long i, const double (&)[W],
                                    const Elem (&)[W], const Con &,
                                    AccT &) const {
    double out[W];
    if (op == 0) {
#pragma unroll
      for (int e = 0; e < W; e++) out[e] = alpha;
    } else if (op == 1) {
      ldv<W>(y, i, out);
#pragma unroll
      for (int e = 0; e < W; e++) out[e] *= alpha;
    } else {
      double xv[W];
      ldv<W>(y, i, out);
      ldv<W>(x, i, xv);
#pragma unroll
      for (int e = 0; e < W; e++) out[e] = fma(alpha, xv[e], out[e]);
    }
    stv<W>(y, i, out);
  }
};

// sum x*y, sum x^2, sum |x|, max |x| in one pass (ParOptVec.cpp:63-143)
struct VecRedF : NoStreams {
  static constexpr int NS = 3, NX = 1, NM = 0, NB = 0;
  typedef Acc<NS, NX, NM> AccT;
  typedef NoElem Elem;
  typedef NoCon Con;
  const double *x, *y;
  template <int W>
  __device__ __forceinline__ void A(long long, const double (&)[W], Elem (&)[W],
                                    double (&)[W][1], AccT *acc) const {}
  __device__ __forceinline__ void B(long long, const double (&)[1], Con &,
                                    AccT &) const {}
  template <int W>
  __device__ __forceinline__ void C(long long i, const double (&)[W],
                                    const Elem (&)[W], const Con &,
                                    AccT &acc) const {
    double xv[W], yv[W];
    ldv<W>(x, i, xv);
    if (y) {
      ldv<W>(y, i, yv);
    } else {
#pragma unroll
      for (int e = 0; e < W; e++) yv[e] = 0.0;
    }
#pragma unroll
    for (int e = 0; e < W; e++) {
      acc.s[0] = fma(xv[e], yv[e], acc.s[0]);
      acc.s[1] = fma(xv[e], xv[e], acc.s[1]);
      acc.s[2] += fabs(xv[e]);
      acc.x[0] = fmax(acc.x[0], fabs(xv[e]));
    }
  }
};

// Multi-vector dot: out[k] = sum_i x[i] * V_k[i] for up to KC vectors per pass
// over x (the reference re-reads x once per vector, ParOptVec.cpp:152-167).
template <int KC>
__global__ void __launch_bounds__(PCU_THREADS)
    mdot_kernel(const double *__restrict__ x, ColTable cols, int col0, int ncols,
                long long n, double *partials, unsigned int *counter,
                double *result) {
  double acc[KC];
#pragma unroll
  for (int k = 0; k < KC; k++) acc[k] = 0.0;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long nvec = n / 2;
  for (long long v = tid; v < nvec; v += nthreads) {
    const double2 xv = *reinterpret_cast<const double2 *>(x + 2 * v);
#pragma unroll
    for (int k = 0; k < KC; k++) {
      if (k < ncols) {
        const double2 c =
            *reinterpret_cast<const double2 *>(cols.p[col0 + k] + 2 * v);
        acc[k] = fma(xv.x, c.x, acc[k]);
        acc[k] = fma(xv.y, c.y, acc[k]);
      }
    }
  }
  if ((n & 1) && tid == 0) {
#pragma unroll
    for (int k = 0; k < KC; k++) {
      if (k < ncols) acc[k] = fma(x[n - 1], cols.p[col0 + k][n - 1], acc[k]);
    }
  }
  // block reduce
  __shared__ double sm[PCU_THREADS / 32][KC];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < KC; k++) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += shfl_down_d(v, o);
    if (lane == 0) sm[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < KC) {
    double v = 0.0;
    for (int w = 0; w < PCU_THREADS / 32; w++) v += sm[w][threadIdx.x];
    partials[(size_t)blockIdx.x * KC + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int k = warp; k < ncols; k += PCU_THREADS / 32) {
      double v = pcu_ordered_sum(partials + k, (size_t)KC, (unsigned)lane, 32u, gridDim.x);
      for (int o = 16; o > 0; o >>= 1) v += shfl_down_d(v, o);
      if (lane == 0) result[col0 + k] = v;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// Enqueue out[k] = x . cols[k], k < ncols, into ctx->d_big[dst_off + k].
int pcu_mdot_enqueue(pcu_ctx *ctx, const double *x, const ColTable &cols,
                     int ncols, long long n, int dst_off) {
  constexpr int KC = 8;
  if (ctx->big_reserve(dst_off + ncols + 8, (size_t)PCU_MAX_BLOCKS * KC)) return 1;
  const int grid = pcu_grid_for(ctx, n);
  for (int c0 = 0; c0 < ncols; c0 += KC) {
    const int nc = (ncols - c0 < KC) ? ncols - c0 : KC;
    ctx->prof_begin("mdot_kernel");
    mdot_kernel<KC><<<grid, PCU_THREADS, 0, ctx->stream>>>(
        x, cols, c0, nc, n, ctx->d_big_partials, ctx->d_counter,
        ctx->d_big + dst_off);
    ctx->prof_end();
    ctx->launches++;
  }
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

// L-SR1 update in one sweep over the stored pairs (QN.cpp:636-747): for pair i
//   Z_i = Y_i - b0 S_i                    (QN.cpp:730-735: the whole Z is rebuilt, b0 is new)
//   s . S_i,  s . Y_i                     (QN.cpp:668-688: the new rows of B and L)
// G pairs per launch: (1 + 3 G) N words instead of the (2 + 3) N per pair of a multi-dot
// over [S | Y] followed by one linear combination per column.  s = the newest stored S.
struct PairTable {
  const double *S[4];
  const double *Y[4];
  double *Z[4];
};
template <int G>
__global__ void __launch_bounds__(PCU_THREADS)
    sr1_pairs_kernel(const double *__restrict__ s, PairTable t, int npairs, double b0,
                     long long n, double *partials, unsigned int *counter, double *result,
                     int dst_s, int dst_y) {
  double acc[2 * G];
#pragma unroll
  for (int k = 0; k < 2 * G; k++) acc[k] = 0.0;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long nvec = n / 2;
  for (long long v = tid; v < nvec; v += nthreads) {
    const double2 sv = *reinterpret_cast<const double2 *>(s + 2 * v);
    double2 a[G], b[G];
#pragma unroll
    for (int g = 0; g < G; g++) {
      if (g < npairs) {
        a[g] = *reinterpret_cast<const double2 *>(t.S[g] + 2 * v);
        b[g] = *reinterpret_cast<const double2 *>(t.Y[g] + 2 * v);
      }
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
      if (g < npairs) {
        double2 z;
        z.x = fma(-b0, a[g].x, b[g].x);
        z.y = fma(-b0, a[g].y, b[g].y);
        *reinterpret_cast<double2 *>(t.Z[g] + 2 * v) = z;
        acc[2 * g] = fma(sv.x, a[g].x, acc[2 * g]);
        acc[2 * g] = fma(sv.y, a[g].y, acc[2 * g]);
        acc[2 * g + 1] = fma(sv.x, b[g].x, acc[2 * g + 1]);
        acc[2 * g + 1] = fma(sv.y, b[g].y, acc[2 * g + 1]);
      }
    }
  }
  if ((n & 1) && tid == 0) {
    const double sv = s[n - 1];
#pragma unroll
    for (int g = 0; g < G; g++) {
      if (g < npairs) {
        const double a = t.S[g][n - 1], b = t.Y[g][n - 1];
        t.Z[g][n - 1] = fma(-b0, a, b);
        acc[2 * g] = fma(sv, a, acc[2 * g]);
        acc[2 * g + 1] = fma(sv, b, acc[2 * g + 1]);
      }
    }
  }
  constexpr int KC = 2 * G;
  __shared__ double sm[PCU_THREADS / 32][KC];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < KC; k++) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += shfl_down_d(v, o);
    if (lane == 0) sm[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < KC) {
    double v = 0.0;
    for (int w = 0; w < PCU_THREADS / 32; w++) v += sm[w][threadIdx.x];
    partials[(size_t)blockIdx.x * KC + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int c = atomicAdd(counter, 1u);
    is_last = (c == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int k = warp; k < 2 * npairs; k += PCU_THREADS / 32) {
      double v = pcu_ordered_sum(partials + k, (size_t)KC, (unsigned)lane, 32u, gridDim.x);
      for (int o = 16; o > 0; o >>= 1) v += shfl_down_d(v, o);
      if (lane == 0) result[((k & 1) ? dst_y : dst_s) + (k >> 1)] = v;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// Enqueue the sweep over `np` pairs: ctx->d_big[i] = s . S_i, ctx->d_big[np + i] = s . Y_i.
int pcu_sr1_pairs_enqueue(pcu_ctx *ctx, const double *s, const double *const *S,
                          const double *const *Y, double *const *Z, int np, double b0,
                          long long n) {
  constexpr int G = 4;
  if (np <= 0) return 0;
  if (ctx->big_reserve(2 * (size_t)np + 8, (size_t)PCU_MAX_BLOCKS * 2 * G)) return 1;
  const int grid = pcu_grid_for(ctx, n);
  for (int p0 = 0; p0 < np; p0 += G) {
    PairTable t;
    const int k = np - p0 < G ? np - p0 : G;
    for (int g = 0; g < G; g++) {
      const int i = p0 + (g < k ? g : 0);
      t.S[g] = S[i];
      t.Y[g] = Y[i];
      t.Z[g] = Z[i];
    }
    ctx->prof_begin("sr1_pairs_kernel");
    sr1_pairs_kernel<G><<<grid, PCU_THREADS, 0, ctx->stream>>>(
        s, t, k, b0, n, ctx->d_big_partials, ctx->d_counter, ctx->d_big, p0, np + p0);
    ctx->prof_end();
    ctx->launches++;
  }
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

template <class F>
static int launch_plain(pcu_ctx *ctx, const F &f, long long n, RedBuf rb) {
  WDesc w;
  memset(&w, 0, sizeof(w));
  return pcu_launch_tile(ctx, f, n, w, rb);
}

static int vec_reduce(pcu_vec *x, pcu_vec *y, double out[4]) {
  pcu_ctx *ctx = x->ctx;
  pcu_vec_ready(x);
  pcu_vec_ready(y);
  VecRedF f;
  f.x = x->d;
  f.y = y ? y->d : nullptr;
  RedBuf rb = ctx->redbuf(3, 1, 0);
  if (launch_plain(ctx, f, x->n, rb)) return 1;
  return ctx->fetch(out);
}

extern "C" {

pcu_vec *pcu_vec_create(pcu_ctx *ctx, int n) {
  if (!ctx || n < 0) return nullptr;
  pcu_vec *v = new pcu_vec;
  v->ctx = ctx;
  v->n = n;
  size_t bytes = ((size_t)n + 2) * sizeof(double);
  v->managed = ctx->managed_vectors != 0;
  const cudaError_t err = v->managed ? cudaMallocManaged(&v->d, bytes) : cudaMalloc(&v->d, bytes);
  if (err != cudaSuccess) {
    fprintf(stderr, "paropt_b200: %s of %zu bytes failed\n",
            v->managed ? "cudaMallocManaged" : "cudaMalloc", bytes);
    delete v;
    return nullptr;
  }
  cudaMemsetAsync(v->d, 0, bytes, ctx->stream);  // first touch on the device
  return v;
}

void pcu_vec_destroy(pcu_vec *v) {
  if (!v) return;
  if (v->owns && v->d) {
    cudaStreamSynchronize(v->ctx->stream);
    cudaFree(v->d);
  }
  delete v;
}

int pcu_vec_size(pcu_vec *v) { return v->n; }

int pcu_vec_set(pcu_vec *v, double alpha) {
  pcu_vec_ready(v);
  VecOpF f;
  f.op = 0;
  f.alpha = alpha;
  f.y = v->d;
  f.x = nullptr;
  RedBuf rb = {nullptr, nullptr, nullptr, 0};
  return launch_plain(v->ctx, f, v->n, rb);
}

int pcu_vec_zero(pcu_vec *v) {
  pcu_vec_ready(v);
  PCU_CUDA_OK(cudaMemsetAsync(v->d, 0, (size_t)v->n * sizeof(double),
                              v->ctx->stream));
  return 0;
}

int pcu_vec_copy(pcu_vec *dst, pcu_vec *src) {
  // a mismatched vector is silently ignored in the reference
  // (ParOptVec.cpp:51-55); here a size mismatch is an error
  if (!src || src->n != dst->n) return 1;
  pcu_vec_ready(dst);
  pcu_vec_ready(src);
  PCU_CUDA_OK(cudaMemcpyAsync(dst->d, src->d, (size_t)dst->n * sizeof(double),
                              cudaMemcpyDeviceToDevice, dst->ctx->stream));
  return 0;
}

int pcu_vec_norm(pcu_vec *v, double *out) {
  double r[4];
  if (vec_reduce(v, nullptr, r)) return 1;
  *out = sqrt(r[1]);
  return 0;
}

int pcu_vec_maxabs(pcu_vec *v, double *out) {
  double r[4];
  if (vec_reduce(v, nullptr, r)) return 1;
  *out = r[3];
  return 0;
}

int pcu_vec_l1norm(pcu_vec *v, double *out) {
  double r[4];
  if (vec_reduce(v, nullptr, r)) return 1;
  *out = r[2];
  return 0;
}

int pcu_vec_dot(pcu_vec *x, pcu_vec *y, double *out) {
  if (!y || y->n != x->n) return 1;
  double r[4];
  if (vec_reduce(x, y, r)) return 1;
  *out = r[0];
  return 0;
}

int pcu_vec_mdot(pcu_vec *x, pcu_vec **vecs, int nvecs, double *out) {
  if (nvecs > PCU_MAX_COLS) return 1;
  ColTable cols;
  for (int k = 0; k < nvecs; k++) {
    if (!vecs[k] || vecs[k]->n != x->n) return 1;
    pcu_vec_ready(vecs[k]);
    cols.p[k] = vecs[k]->d;
  }
  pcu_vec_ready(x);
  if (pcu_mdot_enqueue(x->ctx, x->d, cols, nvecs, x->n, 0)) return 1;
  return x->ctx->big_fetch(nvecs, out);
}

int pcu_vec_scale(pcu_vec *v, double alpha) {
  pcu_vec_ready(v);
  VecOpF f;
  f.op = 1;
  f.alpha = alpha;
  f.y = v->d;
  f.x = nullptr;
  RedBuf rb = {nullptr, nullptr, nullptr, 0};
  return launch_plain(v->ctx, f, v->n, rb);
}

int pcu_vec_axpy(pcu_vec *y, double alpha, pcu_vec *x) {
  if (!x || x->n != y->n) return 1;
  pcu_vec_ready(x);
  pcu_vec_ready(y);
  VecOpF f;
  f.op = 2;
  f.alpha = alpha;
  f.y = y->d;
  f.x = x->d;
  RedBuf rb = {nullptr, nullptr, nullptr, 0};
  return launch_plain(y->ctx, f, y->n, rb);
}

double *pcu_vec_device_ptr(pcu_vec *v) {
  pcu_vec_ready(v);
  return v->d;
}

// ParOptVec::getArray (ParOptVec.cpp:211) for host loops: unified memory, valid on
// the host once the stream has drained; the next library call moves it back.
double *pcu_vec_host_ptr(pcu_vec *v) {
  if (!v || !v->managed) return nullptr;
  if (cudaStreamSynchronize(v->ctx->stream) != cudaSuccess) return nullptr;
  v->host_touched = true;
  return v->d;
}
int pcu_vec_is_managed(pcu_vec *v) { return v && v->managed ? 1 : 0; }

int pcu_vec_to_host(pcu_vec *v, double *host, int n) {
  if (n > v->n) return 1;
  pcu_vec_ready(v);
  PCU_CUDA_OK(cudaMemcpyAsync(host, v->d, (size_t)n * sizeof(double),
                              cudaMemcpyDeviceToHost, v->ctx->stream));
  PCU_CUDA_OK(cudaStreamSynchronize(v->ctx->stream));
  return 0;
}

int pcu_vec_from_host(pcu_vec *v, const double *host, int n) {
  if (n > v->n) return 1;
  pcu_vec_ready(v);
  PCU_CUDA_OK(cudaMemcpyAsync(v->d, host, (size_t)n * sizeof(double),
                              cudaMemcpyHostToDevice, v->ctx->stream));
  PCU_CUDA_OK(cudaStreamSynchronize(v->ctx->stream));
  return 0;
}

}  // extern "C"
