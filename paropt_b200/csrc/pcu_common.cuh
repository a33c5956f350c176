// pcu_common.cuh -- shared device/host helpers of the B200-native ParOpt hot path.
//
// Layout conventions
//   * every distributed vector is one contiguous fp64 device array whose base is
//     256-byte aligned (cudaMalloc), so element pairs (2i, 2i+1) can always be
//     moved with one 128-bit access;
//   * sets of column vectors (the dense constraint gradients A_j and the compact
//     quasi-Newton vectors Z_k) are passed to kernels as a table of base
//     pointers in kernel-parameter (constant) space;
//   * reductions are deterministic: fixed grid, per-thread sequential
//     accumulation, shuffle tree per warp, fixed-order combine of warp and block
//     partials by the last block to finish.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define PCU_MAX_COLS 160   // max ncon + quasi-Newton width handled by kernels
#define PCU_THREADS 256
#define PCU_TILE_THREADS 128  // block size of the fused streaming kernels
#define PCU_MAX_RED 64     // max scalars reduced by one fused kernel
#define PCU_MAX_BLOCKS 4096

#define PCU_CUDA_OK(call)                                                      \
  do {                                                                         \
    cudaError_t err__ = (call);                                                \
    if (err__ != cudaSuccess) {                                                \
      fprintf(stderr, "paropt_b200: CUDA error %s at %s:%d: %s\n",            \
              cudaGetErrorName(err__), __FILE__, __LINE__,                     \
              cudaGetErrorString(err__));                                      \
      return 1;                                                                \
    }                                                                          \
  } while (0)

struct ColTable {
  const double *p[PCU_MAX_COLS];
};
struct CoefTable {
  double v[PCU_MAX_COLS];
};

// Sparse weighting-constraint descriptor as seen by kernels (pcu_weighting).
struct WDesc {
  int nwcon;
  int nw;
  int wstride;
  int mode;          // 0 none, 1 aligned power-of-two blocks (shuffle), 2 generic
  long long wstart;
  long long wend;    // wstart + nwcon * wstride
  double coef0, coef_rest, wconst;
};

// ---------------------------------------------------------------- vector I/O
template <int W>
__device__ __forceinline__ void ldv(const double *__restrict__ p, long long i,
                                    double (&out)[W]);
template <>
__device__ __forceinline__ void ldv<1>(const double *__restrict__ p,
                                       long long i, double (&out)[1]) {
  out[0] = p[i];
}
template <>
__device__ __forceinline__ void ldv<2>(const double *__restrict__ p,
                                       long long i, double (&out)[2]) {
  double2 v = *reinterpret_cast<const double2 *>(p + i);
  out[0] = v.x;
  out[1] = v.y;
}
template <int W>
__device__ __forceinline__ void stv(double *__restrict__ p, long long i,
                                    const double (&in)[W]);
template <>
__device__ __forceinline__ void stv<1>(double *__restrict__ p, long long i,
                                       const double (&in)[1]) {
  p[i] = in[0];
}
template <>
__device__ __forceinline__ void stv<2>(double *__restrict__ p, long long i,
                                       const double (&in)[2]) {
  *reinterpret_cast<double2 *>(p + i) = make_double2(in[0], in[1]);
}

// ---------------------------------------------------------------- reductions
template <int NS, int NX, int NM>
struct Acc {
  double s[NS > 0 ? NS : 1];  // sums
  double x[NX > 0 ? NX : 1];  // maxima
  double m[NM > 0 ? NM : 1];  // minima
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < (NS > 0 ? NS : 1); i++) s[i] = 0.0;
#pragma unroll
    for (int i = 0; i < (NX > 0 ? NX : 1); i++) x[i] = 0.0;
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); i++) m[i] = 1.0e300;
  }
};

// Variant whose sums live in dynamic shared memory ([slot][thread], conflict
// free): for kernels that carry 16-32 running dot products next to 30+ live
// values, registers are what limits the loads in flight.  The functor must
// set SMEM >= NS * PCU_TILE_THREADS * 8.
__device__ __forceinline__ double *pcu_dyn_smem_d() {
  extern __shared__ double2 pcu_dyn_smem[];
  return reinterpret_cast<double *>(pcu_dyn_smem);
}
struct SmemSums {
  double *b;
  __device__ __forceinline__ double &operator[](int j) const {
    return b[j * PCU_TILE_THREADS];
  }
};
template <int NS, int NX, int NM>
struct AccS {
  SmemSums s;
  double x[NX > 0 ? NX : 1];
  double m[NM > 0 ? NM : 1];
  __device__ __forceinline__ void init() {
    s.b = pcu_dyn_smem_d() + threadIdx.x;
#pragma unroll
    for (int i = 0; i < (NS > 0 ? NS : 1); i++) s[i] = 0.0;
#pragma unroll
    for (int i = 0; i < (NX > 0 ? NX : 1); i++) x[i] = 0.0;
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); i++) m[i] = 1.0e300;
  }
};

__device__ __forceinline__ double shfl_xor_d(double v, int o) {
  return __shfl_xor_sync(0xffffffffu, v, o);
}
__device__ __forceinline__ double shfl_down_d(double v, int o) {
  return __shfl_down_sync(0xffffffffu, v, o);
}

struct RedBuf {
  double *partials;        // [gridDim.x][NR]
  unsigned int *counter;   // zero before launch, reset by the last block
  double *result;          // [NR] device
  int prefetch;            // grid-stride iterations to prefetch ahead into L2 (0: off)
};

// Block-level + grid-level deterministic combine.  Every thread of the block
// must call this.  Result layout: sums, then maxima, then minima.
template <int NS, int NX, int NM, class AccT_>
__device__ void finish_reduction(AccT_ &acc, const RedBuf &rb) {
  constexpr int NR = NS + NX + NM;
  __shared__ double sm[PCU_THREADS / 32][NR > 0 ? NR : 1];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NS; i++) {
    double v = acc.s[i];
    for (int o = 16; o > 0; o >>= 1) v += shfl_down_d(v, o);
    if (lane == 0) sm[warp][i] = v;
  }
#pragma unroll
  for (int i = 0; i < NX; i++) {
    double v = acc.x[i];
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, shfl_down_d(v, o));
    if (lane == 0) sm[warp][NS + i] = v;
  }
#pragma unroll
  for (int i = 0; i < NM; i++) {
    double v = acc.m[i];
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, shfl_down_d(v, o));
    if (lane == 0) sm[warp][NS + NX + i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NR) {
    const int i = threadIdx.x;
    double v = sm[0][i];
    for (int w = 1; w < nwarps; w++) {
      if (i < NS) v += sm[w][i];
      else if (i < NS + NX) v = fmax(v, sm[w][i]);
      else v = fmin(v, sm[w][i]);
    }
    rb.partials[(size_t)blockIdx.x * NR + i] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(rb.counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // one warp per value, lanes stride over blocks, fixed order
    for (int i = warp; i < NR; i += nwarps) {
      double v = (i < NS) ? 0.0 : ((i < NS + NX) ? 0.0 : 1.0e300);
      for (unsigned int b = lane; b < gridDim.x; b += 32) {
        double p = rb.partials[(size_t)b * NR + i];
        if (i < NS) v += p;
        else if (i < NS + NX) v = fmax(v, p);
        else v = fmin(v, p);
      }
      for (int o = 16; o > 0; o >>= 1) {
        double q = shfl_down_d(v, o);
        if (i < NS) v += q;
        else if (i < NS + NX) v = fmax(v, q);
        else v = fmin(v, q);
      }
      if (lane == 0) rb.result[i] = v;
    }
    if (threadIdx.x == 0) *rb.counter = 0u;
  }
}

// ------------------------------------------------------- sums of logarithms
// sum_i log(f_i) is accumulated as the logarithm of a running product: the
// mantissa product P in [1, 2) and the exponent sum E live in two accumulator
// slots (P == 0 stands for "empty"), one multiplication and a few integer
// operations per factor instead of one fp64 log.  Factors must be positive and
// in [1e-150, 1e150].  lp_value turns the pair into E ln2 + log(P).
__device__ __forceinline__ void lp_mul(double &P, double &E, double f) {
  double p = (P == 0.0 ? 1.0 : P) * f;
  const long long bits = __double_as_longlong(p);
  const int e = (int)((bits >> 52) & 0x7ff) - 1023;
  P = __longlong_as_double((bits & 0x800FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL);
  E += (double)e;
}
__device__ __forceinline__ double lp_value(double P, double E) {
  return fma(E, 0.693147180559945309417232121458, P == 0.0 ? 0.0 : log(P));
}

// ------------------------------------------------------------ L2 prefetch
// The fused kernels read 10-35 independent streams with 50-130 registers per
// thread, i.e. at 20-50 % occupancy: not enough loads in flight to cover DRAM
// latency.  Instead of buying occupancy, each warp asks the L2 for the 512-byte
// slice of every stream that it will touch in its NEXT grid-stride iteration
// (cp.async.bulk.prefetch.L2, one instruction per stream, spread over the
// lanes): the demand loads then hit in L2 and DRAM runs one iteration ahead.
struct NoStreams {
  static constexpr int NB2 = 0;   // second-round per-constraint block sums (C2 / E phases)
  static constexpr int HASP = 0;  // P(ci, con): per-constraint prologue seen by A (AP form)
  static constexpr int NF = 0;    // third round F / FG after E: E leaves con.d[FD] to broadcast
  static constexpr int FD = 0;
  static constexpr int SMEM = 0;  // dynamic shared memory per block (bytes)
  template <class A_>
  __device__ __forceinline__ void finalize(A_ &) const {}  // per-thread, before the combine
  static constexpr int MINB = 7;  // __launch_bounds__ minimum blocks per SM (<= 73 regs)
  template <class P>
  __device__ __forceinline__ void streams(P &) const {}
};
struct Prefetcher {
  long long off;  // element offset of the warp's next 64-element slice
  int lane;
  int idx;
  __device__ __forceinline__ void operator()(const double *ptr) {
    if (ptr != nullptr && ((idx++) & 31) == lane) {
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr + off),
                   "r"(512)
                   : "memory");
    }
  }
};

struct PrefetcherNow {  // same-iteration prefetch of the thread's own 16 bytes
  long long i;
  __device__ __forceinline__ void operator()(const double *ptr) {
    if (ptr != nullptr)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + i) : "memory");
  }
};

// ------------------------------------------------------------ tile harness
// A fused kernel is a functor F with
//   static constexpr int NS, NX, NM   number of sum / max / min accumulators
//   static constexpr int NB           number of per-constraint block sums (0..2)
//   struct Elem, struct Con           per-element / per-constraint registers
//   template<int W> void A(i, coef[W], Elem[W], part[W][NB]) const  -- loads the
//        element data and returns the terms of the block sums (idempotent)
//   void B(ci, sum[NB], Con&, acc) const -- ONE thread per weighting constraint;
//        Con::d[ND] is then broadcast to the constraint's elements
//   template<int W> void C(i, coef[W], Elem[W], Con, acc) const -- finishes the
//        elements (coef = 0 and Con = zero outside weighting constraints)
template <class F>
__device__ __forceinline__ void generic_range(const F &f, const WDesc &w,
                                              long long lo, long long hi,
                                              long long tid, long long nthreads,
                                              typename F::AccT &acc) {
  constexpr int NB = F::NB > 0 ? F::NB : 1;
  // (1) elements outside every weighting constraint
  for (long long i = lo + tid; i < hi; i += nthreads) {
    bool in_con = false;
    if (w.nwcon > 0 && i >= w.wstart && i < w.wend) {
      in_con = ((i - w.wstart) % w.wstride) < w.nw;
    }
    if (!in_con) {
      typename F::Elem e[1];
      double coef[1] = {0.0};
      double part[1][NB];
      if constexpr (F::HASP) { typename F::Con c0; c0.zero(); f.template AP<1>(i, coef, e, part, &acc, c0); } else { f.template A<1>(i, coef, e, part, &acc); }
      typename F::Con con;
      con.zero();
      if constexpr (F::NB2 > 0) {
        double part2[1][F::NB2];
        f.template C2<1>(i, coef, e, con, acc, part2);
      } else {
        f.template C<1>(i, coef, e, con, acc);
      }
      if constexpr (F::NF > 0) f.FG(i, coef[0], con, acc);
    }
  }
  // (2) whole constraints, one thread per constraint
  if (w.nwcon > 0) {
    long long c_lo = 0, c_hi = w.nwcon;
    if (lo > w.wstart) c_lo = (lo - w.wstart + w.wstride - 1) / w.wstride;
    if (hi < w.wend) c_hi = (hi - w.wstart) / w.wstride;
    for (long long ci = c_lo + tid; ci < c_hi; ci += nthreads) {
      const long long j0 = w.wstart + ci * w.wstride;
      double sum[NB];
#pragma unroll
      for (int b = 0; b < NB; b++) sum[b] = 0.0;
      typename F::Con con;
      con.zero();
      if constexpr (F::HASP) f.P(ci, con);
      for (int k = 0; k < w.nw; k++) {
        typename F::Elem e[1];
        double coef[1] = {k == 0 ? w.coef0 : w.coef_rest};
        double part[1][NB];
        if constexpr (F::HASP) { f.template AP<1>(j0 + k, coef, e, part, (typename F::AccT *)nullptr, con); } else { f.template A<1>(j0 + k, coef, e, part, (typename F::AccT *)nullptr); }
#pragma unroll
        for (int b = 0; b < NB; b++) sum[b] += part[0][b];
      }
      f.B(ci, sum, con, acc);
      double sum2[F::NB2 > 0 ? F::NB2 : 1];
#pragma unroll
      for (int b = 0; b < (F::NB2 > 0 ? F::NB2 : 1); b++) sum2[b] = 0.0;
      for (int k = 0; k < w.nw; k++) {
        typename F::Elem e[1];
        double coef[1] = {k == 0 ? w.coef0 : w.coef_rest};
        double part[1][NB];
        if constexpr (F::HASP) { f.template AP<1>(j0 + k, coef, e, part, &acc, con); } else { f.template A<1>(j0 + k, coef, e, part, &acc); }
        if constexpr (F::NB2 > 0) {
          double part2[1][F::NB2];
          f.template C2<1>(j0 + k, coef, e, con, acc, part2);
#pragma unroll
          for (int b = 0; b < F::NB2; b++) sum2[b] += part2[0][b];
        } else {
          f.template C<1>(j0 + k, coef, e, con, acc);
        }
      }
      if constexpr (F::NB2 > 0) f.E(ci, sum2, con, acc);
      if constexpr (F::NF > 0) {
        for (int k = 0; k < w.nw; k++)
          f.FG(j0 + k, k == 0 ? w.coef0 : w.coef_rest, con, acc);
      }
    }
  }
}

template <class F>
__global__ void __launch_bounds__(PCU_TILE_THREADS, F::MINB)
    tile_kernel(const F f, const long long n, const WDesc w, const RedBuf rb) {
  constexpr int NB = F::NB > 0 ? F::NB : 1;
  typename F::AccT acc;
  acc.init();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;

  if (w.mode == 2) {
    generic_range(f, w, 0, n, tid, nthreads, acc);
  } else {
    // 128-bit path over full warps of element pairs; the ragged tail (< 64
    // elements + whole constraints) goes through the generic path.
    const long long nvec_main = ((n / 2) / 32) * 32;
    const long long ncon_elems = (w.mode == 1) ? (long long)w.nwcon * w.nw : 0;
    const int half = (w.mode == 1) ? (w.nw >> 1) : 1;  // lanes per constraint
    for (long long v = tid; v < nvec_main; v += nthreads) {
      const long long i = 2 * v;
      if (rb.prefetch < 0) {
        PrefetcherNow pn;
        pn.i = i;
        f.streams(pn);
      }
      if (rb.prefetch > 0 && v + rb.prefetch * nthreads < nvec_main) {
        Prefetcher pf;
        pf.lane = threadIdx.x & 31;
        pf.off = 2 * (v + rb.prefetch * nthreads - pf.lane);
        pf.idx = 0;
        f.streams(pf);
      }
      typename F::Elem e[2];
      double coef[2] = {0.0, 0.0};
      double part[2][NB];
      const bool in_con = i < ncon_elems;
      if (in_con) {
        const int k = (int)(i & (long long)(w.nw - 1));
        coef[0] = (k == 0) ? w.coef0 : w.coef_rest;
        coef[1] = w.coef_rest;
      }
      typename F::Con con;
      con.zero();
      if constexpr (F::HASP) {
        if (in_con) f.P(i / w.nw, con);
        f.template AP<2>(i, coef, e, part, &acc, con);
      } else {
        f.template A<2>(i, coef, e, part, &acc);
      }
      if (w.mode == 1) {
        double sum[NB];
#pragma unroll
        for (int b = 0; b < NB; b++) sum[b] = in_con ? part[0][b] + part[1][b] : 0.0;
        if (F::NB > 0) {
          for (int o = 1; o < half; o <<= 1) {
#pragma unroll
            for (int b = 0; b < NB; b++) sum[b] += shfl_xor_d(sum[b], o);
          }
        }
        const int lane = threadIdx.x & 31;
        const int lead = lane & ~(half - 1);
        if (in_con && lane == lead) f.B(i / w.nw, sum, con, acc);
        if (F::Con::ND > 0) {
#pragma unroll
          for (int b = 0; b < (F::Con::ND > 0 ? F::Con::ND : 1); b++)
            con.d[b] = __shfl_sync(0xffffffffu, con.d[b], lead);
        }
      }
      if constexpr (F::NB2 > 0) {
        double part2[2][F::NB2];
        f.template C2<2>(i, coef, e, con, acc, part2);
        if (w.mode == 1) {
          double sum2[F::NB2];
#pragma unroll
          for (int b = 0; b < F::NB2; b++)
            sum2[b] = in_con ? part2[0][b] + part2[1][b] : 0.0;
          for (int o = 1; o < half; o <<= 1) {
#pragma unroll
            for (int b = 0; b < F::NB2; b++) sum2[b] += shfl_xor_d(sum2[b], o);
          }
          const int lane = threadIdx.x & 31;
          const int lead = lane & ~(half - 1);
          if (in_con && lane == lead) f.E(i / w.nw, sum2, con, acc);
          if constexpr (F::NF > 0)
            con.d[F::FD] = __shfl_sync(0xffffffffu, con.d[F::FD], lead);
        }
        if constexpr (F::NF > 0) f.template F<2>(i, coef, e, con, acc);
      } else {
        f.template C<2>(i, coef, e, con, acc);
      }
    }
    const long long tail_lo = 2 * nvec_main;
    if (tail_lo < n && blockIdx.x == 0) {
      WDesc wt = w;
      if (w.mode == 0) wt.nwcon = 0;
      generic_range(f, wt, tail_lo, n, threadIdx.x, blockDim.x, acc);
    }
  }
  if (F::NS + F::NX + F::NM > 0) {
    f.finalize(acc);
    finish_reduction<F::NS, F::NX, F::NM>(acc, rb);
  }
}

// x + a*p clipped into [lb + dp, ub - dp]  (computeStep, IP.cpp:3148-3191).
// One definition so that the line-search trial point and the accepted point are
// bit-identical.
__device__ __forceinline__ double step_clip(double x, double a, double p,
                                            double lb, double ub, double dp) {
  double r = fma(a, p, x);
  if (r <= lb + dp) r = lb + dp;
  if (r + dp >= ub) r = ub - dp;
  return r;
}
__device__ __forceinline__ double step_clip0(double x, double a, double p,
                                             double dp) {
  double r = fma(a, p, x);
  if (r <= 0.0 + dp) r = 0.0 + dp;
  return r;
}
