// stream_probe.cu -- what HBM bandwidth does a many-stream fp64 pass reach on
// this GPU, as a function of the access scheme?  (Design probe for the fused
// KKT passes: ~34 input vectors, 3 updated in place.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
//   ./stream_probe [K inputs] [J in-place outputs] [log2 n]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define MAXK 48
struct Tab {
  double *p[MAXK];
};

// ---------------------------------------------------------------- register fed
template <int W, int PF>
__global__ void __launch_bounds__(128) reg_kernel(Tab t, int K, int J, long long n) {
  const long long nth = (long long)gridDim.x * blockDim.x;
  const long long nv = n / 2;
  if (W == 1) {
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += nth) {
      if (PF) {
        for (int k = 0; k < K; k++)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(t.p[k] + 2 * v) : "memory");
      }
      double2 s = make_double2(0.0, 0.0);
      for (int k = 0; k < K; k++) {
        const double2 a = *reinterpret_cast<const double2 *>(t.p[k] + 2 * v);
        s.x = fma(a.x, 1.0 + k, s.x);
        s.y = fma(a.y, 1.0 + k, s.y);
      }
      for (int j = 0; j < J; j++) {
        double2 o = make_double2(s.x + j, s.y - j);
        *reinterpret_cast<double2 *>(t.p[j] + 2 * v) = o;
      }
    }
  } else {
    // each warp owns W consecutive 512-byte slices per stream
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = nth >> 5;
    const long long nchunk = nv / (32 * W);
    for (long long c = warp; c < nchunk; c += nwarps) {
      const long long v0 = c * 32 * W + lane;
      if (PF) {
        for (int k = 0; k < K; k++)
#pragma unroll
          for (int w = 0; w < W; w++)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(t.p[k] + 2 * (v0 + 32 * w)) : "memory");
      }
      double2 s[W];
#pragma unroll
      for (int w = 0; w < W; w++) s[w] = make_double2(0.0, 0.0);
      for (int k = 0; k < K; k++) {
#pragma unroll
        for (int w = 0; w < W; w++) {
          const double2 a = *reinterpret_cast<const double2 *>(t.p[k] + 2 * (v0 + 32 * w));
          s[w].x = fma(a.x, 1.0 + k, s[w].x);
          s[w].y = fma(a.y, 1.0 + k, s[w].y);
        }
      }
      for (int j = 0; j < J; j++)
#pragma unroll
        for (int w = 0; w < W; w++) {
          double2 o = make_double2(s[w].x + j, s[w].y - j);
          *reinterpret_cast<double2 *>(t.p[j] + 2 * (v0 + 32 * w)) = o;
        }
    }
  }
}

// ---------------------------------------------------------------- bulk-copy fed
__device__ __forceinline__ unsigned su32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes,
                                         unsigned bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(src), "r"(bytes)
               : "memory");
}

// NCW consumer warps + 1 producer warp; a tile = ROWS rows of every stream.
// OUTMODE 0: consumers store straight to global; 1: through shared memory + bulk store.
template <int OUTMODE>
__global__ void __launch_bounds__(1024, 1)
    tma_kernel(Tab t, int K, int J, long long n, int rows, int nstages, int ncw) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[8], empty[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned colb = rows * 8;
  const unsigned stage_bytes = colb * (K + (OUTMODE ? J : 0));
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; s++) {
      mbar_init(su32(&full[s]), 1);
      mbar_init(su32(&empty[s]), ncw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long ntiles = n / rows;
  if (warp == ncw) {
    long long it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
      const int s = (int)(it % nstages);
      const unsigned round = (unsigned)(it / nstages);
      if (round > 0) mbar_wait(su32(&empty[s]), (round - 1) & 1);
      const unsigned fb = su32(&full[s]);
      if (lane == 0) mbar_expect_tx(fb, colb * K);
      __syncwarp();
      const unsigned base = su32(smem + (size_t)s * stage_bytes);
      for (int k = lane; k < K; k += 32) bulk_g2s(base + k * colb, t.p[k] + tile * rows, colb, fb);
    }
  } else if (warp < ncw) {
    long long it = 0;
    const int nthr = ncw * 32;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
      const int s = (int)(it % nstages);
      mbar_wait(su32(&full[s]), (unsigned)(it / nstages) & 1);
      unsigned char *base = smem + (size_t)s * stage_bytes;
      for (int v = threadIdx.x; v < rows / 2; v += nthr) {
        double2 sacc = make_double2(0.0, 0.0);
        for (int k = 0; k < K; k++) {
          const double2 a = *reinterpret_cast<const double2 *>(base + k * colb + v * 16);
          sacc.x = fma(a.x, 1.0 + k, sacc.x);
          sacc.y = fma(a.y, 1.0 + k, sacc.y);
        }
        for (int j = 0; j < J; j++) {
          double2 o = make_double2(sacc.x + j, sacc.y - j);
          if (OUTMODE == 0)
            *reinterpret_cast<double2 *>(t.p[j] + tile * rows + 2 * v) = o;
          else
            *reinterpret_cast<double2 *>(base + (K + j) * colb + v * 16) = o;
        }
      }
      if (OUTMODE == 1) {
        // consumers -> one elected thread issues the bulk stores, then the stage is released
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
        if (threadIdx.x == 0) {
          for (int j = 0; j < J; j++)
            bulk_s2g(t.p[j] + tile * rows, su32(base + (K + j) * colb), colb);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(su32(&empty[s]));
    }
  }
}

// NP producer warps (streams dealt round robin over producer lanes).
__global__ void __launch_bounds__(1024, 1)
    tmap_kernel(Tab t, int K, int J, long long n, int rows, int nstages, int ncw, int np) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[16], empty[16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned colb = rows * 8;
  const unsigned stage_bytes = colb * K;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; s++) {
      mbar_init(su32(&full[s]), np);
      mbar_init(su32(&empty[s]), rows / 64);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long ntiles = n / rows;
  if (warp >= ncw) {
    const int pw = warp - ncw;
    int mine = 0;
    for (int k = pw * 32 + lane; k < K; k += 32 * np) mine++;
    unsigned tot = 0;
    for (int o = 0; o < 32; o++) tot += __shfl_sync(0xffffffffu, mine, o);
    long long it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
      const int s = (int)(it % nstages);
      const unsigned round = (unsigned)(it / nstages);
      if (round > 0) mbar_wait(su32(&empty[s]), (round - 1) & 1);
      const unsigned fb = su32(&full[s]);
      if (lane == 0) mbar_expect_tx(fb, colb * tot);
      __syncwarp();
      const unsigned base = su32(smem + (size_t)s * stage_bytes);
      for (int k = pw * 32 + lane; k < K; k += 32 * np)
        bulk_g2s(base + k * colb, t.p[k] + tile * rows, colb, fb);
    }
  } else {
    // consumer group g = warps [g*wpt, (g+1)*wpt) takes tiles it % ngroups == g
    const int wpt = rows / 64;           // warps per tile
    const int ngroups = ncw / wpt;
    const int g = warp / wpt, wg = warp % wpt;
    long long it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
      if ((int)(it % ngroups) != g) continue;
      const int s = (int)(it % nstages);
      mbar_wait(su32(&full[s]), (unsigned)(it / nstages) & 1);
      unsigned char *base = smem + (size_t)s * stage_bytes;
      const int v = wg * 32 + lane;
      double2 sacc = make_double2(0.0, 0.0);
      for (int k = 0; k < K; k++) {
        const double2 a = *reinterpret_cast<const double2 *>(base + k * colb + v * 16);
        sacc.x = fma(a.x, 1.0 + k, sacc.x);
        sacc.y = fma(a.y, 1.0 + k, sacc.y);
      }
      for (int j = 0; j < J; j++) {
        double2 o = make_double2(sacc.x + j, sacc.y - j);
        *reinterpret_cast<double2 *>(t.p[j] + tile * rows + 2 * v) = o;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(su32(&empty[s]));
    }
  }
}

// warp-private rings: every warp streams its own 64-row tiles, no producer warp
__global__ void __launch_bounds__(1024, 1)
    wp_kernel(Tab t, int K, int J, long long n, int nstages) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[32 * 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned colb = 512;
  const unsigned stage_bytes = colb * K;
  unsigned char *mybase = smem + (size_t)warp * nstages * stage_bytes;
  if (lane == 0)
    for (int s = 0; s < nstages; s++) mbar_init(su32(&full[warp * 8 + s]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long ntiles = n / 64;
  const long long gw = (long long)blockIdx.x * nw + warp, ngw = (long long)gridDim.x * nw;
  auto issue = [&](long long tile, int s) {
    const unsigned fb = su32(&full[warp * 8 + s]);
    if (lane == 0) mbar_expect_tx(fb, colb * K);
    __syncwarp();
    const unsigned base = su32(mybase + (size_t)s * stage_bytes);
    for (int k = lane; k < K; k += 32) bulk_g2s(base + k * colb, t.p[k] + tile * 64, colb, fb);
  };
  long long tile = gw;
  for (int s = 0; s < nstages - 1; s++)
    if (tile + s * ngw < ntiles) issue(tile + s * ngw, s);
  long long it = 0;
  for (; tile < ntiles; tile += ngw, it++) {
    const int s = (int)(it % nstages);
    const long long nxt = tile + (long long)(nstages - 1) * ngw;
    if (nxt < ntiles) issue(nxt, (int)((it + nstages - 1) % nstages));
    mbar_wait(su32(&full[warp * 8 + s]), (unsigned)(it / nstages) & 1);
    unsigned char *base = mybase + (size_t)s * stage_bytes;
    double2 sacc = make_double2(0.0, 0.0);
    for (int k = 0; k < K; k++) {
      const double2 a = *reinterpret_cast<const double2 *>(base + k * colb + lane * 16);
      sacc.x = fma(a.x, 1.0 + k, sacc.x);
      sacc.y = fma(a.y, 1.0 + k, sacc.y);
    }
    for (int j = 0; j < J; j++) {
      double2 o = make_double2(sacc.x + j, sacc.y - j);
      *reinterpret_cast<double2 *>(t.p[j] + tile * 64 + 2 * lane) = o;
    }
    __syncwarp();
  }
}

static float time_it(void (*launch)(void *), void *arg, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  launch(arg);
  launch(arg);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < reps; r++) launch(arg);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

struct Args {
  Tab t;
  int K, J;
  long long n;
  int variant, bps, rows, stages, ncw, np;
};

static void launch(void *p) {
  Args *a = (Args *)p;
  const int grid = 148 * a->bps;
  switch (a->variant) {
    case 0: reg_kernel<1, 0><<<grid, 128>>>(a->t, a->K, a->J, a->n); break;
    case 1: reg_kernel<1, 1><<<grid, 128>>>(a->t, a->K, a->J, a->n); break;
    case 2: reg_kernel<2, 1><<<grid, 128>>>(a->t, a->K, a->J, a->n); break;
    case 3: reg_kernel<4, 1><<<grid, 128>>>(a->t, a->K, a->J, a->n); break;
    case 4: reg_kernel<2, 0><<<grid, 128>>>(a->t, a->K, a->J, a->n); break;
    case 20: {
      const size_t smem = (size_t)a->stages * a->rows * 8 * a->K;
      cudaFuncSetAttribute(tmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      tmap_kernel<<<148, 32 * (a->ncw + a->np), smem>>>(a->t, a->K, a->J, a->n, a->rows,
                                                        a->stages, a->ncw, a->np);
      break;
    }
    case 21: {
      const size_t smem = (size_t)a->stages * 512 * a->K * a->ncw;
      cudaFuncSetAttribute(wp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      wp_kernel<<<148, 32 * a->ncw, smem>>>(a->t, a->K, a->J, a->n, a->stages);
      break;
    }
    case 10:
    case 11: {
      const int outm = a->variant - 10;
      const size_t smem = (size_t)a->stages * a->rows * 8 * (a->K + (outm ? a->J : 0));
      auto kern = outm ? tma_kernel<1> : tma_kernel<0>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<148 * a->bps, 32 * (a->ncw + 1), smem>>>(a->t, a->K, a->J, a->n, a->rows,
                                                       a->stages, a->ncw);
      break;
    }
  }
}

int main(int argc, char **argv) {
  setvbuf(stdout, NULL, _IONBF, 0);
  Args a;
  a.K = argc > 1 ? atoi(argv[1]) : 34;
  a.J = argc > 2 ? atoi(argv[2]) : 3;
  const int lg = argc > 3 ? atoi(argv[3]) : 26;
  a.n = 1LL << lg;
  for (int k = 0; k < a.K; k++) {
    if (cudaMalloc(&a.t.p[k], a.n * 8) != cudaSuccess) {
      printf("alloc failed\n");
      return 1;
    }
    cudaMemset(a.t.p[k], 0, a.n * 8);
  }
  const double gb = (double)(a.K + a.J) * a.n * 8 / 1e9;
  printf("K=%d J=%d n=2^%d  bytes/launch %.2f GB\n", a.K, a.J, lg, gb);
  const char *names[] = {"reg W=1", "reg W=1 +pf", "reg W=2 +pf", "reg W=4 +pf", "reg W=2"};
  const int quick = argc > 4 ? atoi(argv[4]) : 0;
  for (int v = 0; v < (quick ? 0 : 5); v++) {
    for (int bps : {4, 6, 8, 12, 16}) {
      a.variant = v;
      a.bps = bps;
      const float ms = time_it(launch, &a, 5);
      printf("%-14s blocks/SM %2d : %7.3f ms  %7.1f GB/s\n", names[v], bps, ms, gb / ms * 1e3);
    }
  }
  for (int outm = 0; outm < (quick ? 0 : 2); outm++)
    for (int rows : {128, 256, 512})
      for (int ncw : {4, 8, 16})
        for (int bps : {1, 2}) {
          const size_t col = (size_t)rows * 8 * (a.K + (outm ? a.J : 0));
          int stages = (int)((220 * 1024 / bps) / col);
          if (stages > 8) stages = 8;
          if (stages < 2) continue;
          a.variant = 10 + outm;
          a.rows = rows;
          a.ncw = ncw;
          a.stages = stages;
          a.bps = bps;
          const float ms = time_it(launch, &a, 5);
          cudaError_t e = cudaGetLastError();
          printf("tma out=%d rows %3d ncw %2d cta/SM %d stages %d : %7.3f ms  %7.1f GB/s %s\n", outm,
                 rows, ncw, bps, stages, ms, gb / ms * 1e3,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  // G consumer groups of rows/64 warps; nstages = G * depth so that a stage always belongs to one group
  for (int rows : {64, 128, 256})
    for (int np : {1, 2, 4, 6})
      for (int depth : {1, 2}) {
        const size_t col = (size_t)rows * 8 * a.K;
        int stages = (int)((216 * 1024) / col);
        int G = stages / depth;
        const int wpt = rows / 64;
        if (G * wpt + np > 30) G = (30 - np) / wpt;
        if (G > 16 / depth) G = 16 / depth;
        if (G < 1) continue;
        stages = G * depth;
        a.variant = 20; a.rows = rows; a.ncw = G * wpt; a.stages = stages; a.np = np;
        const float ms = time_it(launch, &a, 5);
        cudaError_t e = cudaGetLastError();
        printf("tmap rows %3d np %d groups %2d depth %d (ncw %2d stages %2d) : %7.3f ms  %7.1f GB/s %s\n",
               rows, np, G, depth, a.ncw, stages, ms, gb / ms * 1e3,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  for (int ncw : {4, 6, 8, 12})
    for (int stages : {2, 3}) {
      if ((size_t)stages * 512 * a.K * ncw > 220 * 1024) continue;
      a.variant = 21; a.ncw = ncw; a.stages = stages;
      const float ms = time_it(launch, &a, 5);
      cudaError_t e = cudaGetLastError();
      printf("warp-private ncw %2d stages %d : %7.3f ms  %7.1f GB/s %s\n", ncw, stages, ms,
             gb / ms * 1e3, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
