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
  int nw_log2;       // mode 1: nw is a power of two (constraint of element i: i >> nw_log2)
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

// ---------------------------------------------------------------- division
// Branch-free fp64 quotient for the fused passes: hardware reciprocal seed
// (2^-23), two Newton steps, one residual correction -- 9 dependent-free-to-
// interleave instructions instead of the IEEE sequence with its slow-path
// branch (which ends a basic block per quotient and serialises the 6-12
// quotients of a tile).  Operands here are positive, normal and far from the
// range ends (distances to bounds, slacks, multipliers); the result is within
// one ulp of the correctly rounded quotient.
#ifndef PCU_FAST_DIV
#define PCU_FAST_DIV 1
#endif
__device__ __forceinline__ double pcu_rcp(double b) {
#if PCU_FAST_DIV
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / b;
#endif
}
__device__ __forceinline__ double pcu_div(double a, double b) {
#if PCU_FAST_DIV
  const double r = pcu_rcp(b);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
#else
  return a / b;
#endif
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

// Last-block combines: sum of p[b * stride], b = first, first + step, ... < nb, added in
// that order, with the (independent) loads issued sixteen at a time -- a plain
// `for (b) v += p[...]` waits one L2 round trip per block (148 CTAs: ~11 us in the Gram
// kernels, ~2 us per value in finish_reduction; launch_overhead_probe.py).  ld.cg: the
// partials were written by other SMs.
__device__ __forceinline__ double pcu_ordered_sum(const double *p, size_t stride,
                                                  unsigned first, unsigned step, unsigned nb) {
  double v = 0.0;
  unsigned b = first;
  for (; b + 15u * step < nb; b += 16u * step) {
    double t[16];
#pragma unroll
    for (int u = 0; u < 16; u++) t[u] = __ldcg(p + (size_t)(b + (unsigned)u * step) * stride);
#pragma unroll
    for (int u = 0; u < 16; u++) v += t[u];
  }
  double t[16];
#pragma unroll
  for (int u = 0; u < 16; u++) {
    const unsigned bb = b + (unsigned)u * step;
    t[u] = bb < nb ? __ldcg(p + (size_t)bb * stride) : 0.0;
  }
#pragma unroll
  for (int u = 0; u < 16; u++) {
    if (b + (unsigned)u * step < nb) v += t[u];
  }
  return v;
}

struct RedBuf {
  double *partials;        // [gridDim.x][NR]
  unsigned int *counter;   // zero before launch, reset by the last block
  double *result;          // [NR] device
  int prefetch;            // grid-stride iterations to prefetch ahead into L2 (0: off)
  // Zero-copy hand-over to the host (optional): the last block also stores the NR
  // results into page-locked host memory mapped into the device address space and
  // then publishes `seq` in *hflag (system-scope release); the host polls the flag
  // instead of enqueueing a copy and synchronising the stream.
  double *hres;
  unsigned long long *hflag;
  unsigned long long seq;
};

// Block-level + grid-level deterministic combine.  Every thread of the block
// must call this.  Result layout: sums, then maxima, then minima.
template <int NS, int NX, int NM, class AccT_, int MAXW = PCU_THREADS / 32>
__device__ void finish_reduction(AccT_ &acc, const RedBuf &rb, const int tid_ = -1,
                                 const int nthr_ = 0) {
  // default: every thread of the block; otherwise the nthr_ threads (whole warps)
  // whose index within the group is tid_, synchronised on named barrier 1
  const bool part = tid_ >= 0;
  const int tid = part ? tid_ : (int)threadIdx.x;
  const int nthr = part ? nthr_ : (int)blockDim.x;
#define PCU_RED_SYNC()                                              \
  do {                                                              \
    if (part) asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory"); \
    else __syncthreads();                                           \
  } while (0)
  constexpr int NR = NS + NX + NM;
  __shared__ double sm[MAXW][NR > 0 ? NR : 1];
  __shared__ bool is_last;
  const int lane = tid & 31, warp = tid >> 5;
  const int nwarps = nthr >> 5;
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
  PCU_RED_SYNC();
  if (tid < NR) {
    const int i = tid;
    double v = sm[0][i];
    for (int w = 1; w < nwarps; w++) {
      if (i < NS) v += sm[w][i];
      else if (i < NS + NX) v = fmax(v, sm[w][i]);
      else v = fmin(v, sm[w][i]);
    }
    rb.partials[(size_t)blockIdx.x * NR + i] = v;
  }
  __threadfence();
  PCU_RED_SYNC();
  if (tid == 0) {
    unsigned int t = atomicAdd(rb.counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  PCU_RED_SYNC();
  if (is_last) {
    __threadfence();
    // one warp per value, lanes stride over blocks, fixed order
    for (int i = warp; i < NR; i += nwarps) {
      const double ident = (i < NS) ? 0.0 : ((i < NS + NX) ? 0.0 : 1.0e300);
      double v = ident;
      // eight independent loads in flight per lane (one L2 round trip per batch
      // instead of one per block), combined in the order b = lane, lane + 32, ...
      for (unsigned int b0 = lane; b0 < gridDim.x; b0 += 256u) {
        double p[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const unsigned int b = b0 + 32u * (unsigned)u;
          p[u] = b < gridDim.x ? __ldcg(rb.partials + (size_t)b * NR + i) : ident;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          if (b0 + 32u * (unsigned)u < gridDim.x) {
            if (i < NS) v += p[u];
            else if (i < NS + NX) v = fmax(v, p[u]);
            else v = fmin(v, p[u]);
          }
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        double q = shfl_down_d(v, o);
        if (i < NS) v += q;
        else if (i < NS + NX) v = fmax(v, q);
        else v = fmin(v, q);
      }
      if (lane == 0) {
        rb.result[i] = v;
        sm[0][i] = v;
      }
    }
    if (tid == 0) *rb.counter = 0u;
    if (rb.hres) {  // (uniform) publish: ONE thread stores every value, fences once
      // (system scope: one PCIe round trip instead of one per value -- 4-5 us of every
      // fused kernel, scripts/reduce_probe.py) and raises the flag
      PCU_RED_SYNC();
      if (tid == 0) {
#pragma unroll 1
        for (int i = 0; i < NR; i++) rb.hres[i] = sm[0][i];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(rb.hflag) = rb.seq;
      }
    }
  }
#undef PCU_RED_SYNC
}

// ------------------------------------------------------- sums of logarithms
// sum_i log(f_i) is accumulated as the logarithm of a running product: the
// mantissa product P in [1, 2) and the exponent sum E live in two accumulator
// slots (P == 0 stands for "empty"), one multiplication and a few integer
// operations per factor instead of one fp64 log.  lp_value turns the pair into
// E ln2 + log(P).  The fast path needs a positive, normal factor in
// [1e-150, 1e150]; anything else -- zero (a variable ON its bound: lb + dp == lb
// for |lb| >~ 100 or design_precision = 0), negative, NaN, a product that left
// the range -- takes log2(f) itself into E, so that log(0) = -inf (the reference's
// merit becomes +inf and the line search rejects the trial, IP.cpp:3541-3590) and
// NaN propagate exactly as a plain sum of logarithms would.
__device__ __forceinline__ void lp_mul(double &P, double &E, double f) {
  if (!(f >= 1e-150 && f <= 1e150)) {
    E += log(f) * 1.44269504088896340735992468100189214;
    return;
  }
  double p = (P == 0.0 ? 1.0 : P) * f;
  const long long bits = __double_as_longlong(p);
  const int e = (int)((bits >> 52) & 0x7ff) - 1023;
  P = __longlong_as_double((bits & 0x800FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL);
  E += (double)e;
}
// Two factors whose product may leave the fast range (bounds near max_bound_value).
__device__ __forceinline__ void lp_mul2(double &P, double &E, double f0, double f1) {
  const double f = f0 * f1;
  if (f >= 1e-150 && f <= 1e150) {
    lp_mul(P, E, f);
  } else {
    lp_mul(P, E, f0);
    lp_mul(P, E, f1);
  }
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
// Where a functor's loads come from.  Functors with SRC = 1 take a source as
// first argument of every phase and name a slot with each load: GSrc reads the
// global arrays (register-fed tile_kernel, ragged tails), SSrc<ROWS> reads the
// shared-memory stage the bulk-copy engine filled (tma_tile_kernel).
struct GSrc {
  template <int W>
  __device__ __forceinline__ void ld(int, const double *p, long long i,
                                     double (&out)[W]) const {
    ldv<W>(p, i, out);
  }
  template <int W>  // column j of the functor's column table (slot NFIX + j)
  __device__ __forceinline__ void ldc(int, const double *p, long long i,
                                      double (&out)[W]) const {
    ldv<W>(p, i, out);
  }
  __device__ __forceinline__ double ldw(int, const double *p, long long ci) const {
    return p[ci];
  }
  template <int W>  // scratch slots exist in the staged sources only
  __device__ __forceinline__ void st(int, long long, const double (&)[W]) const {
    __trap();
  }
};
template <int ROWS, int NFIX>
struct SSrc {
  unsigned nb;       // shared-space address of the stage: N-slots [compact slot][ROWS] doubles
  unsigned wb;       // ... of its W-slots: [slot][wpitch bytes]
  long long row0, con0;  // first element / first weighting constraint of the tile
  int wpitch;
  unsigned off[NFIX > 0 ? NFIX : 1];  // byte offset of each fixed slot id (from the launch plan)
  unsigned col0;                      // byte offset of column 0 (slot id NFIX)
  template <int W>
  __device__ __forceinline__ void lds(unsigned a, double (&out)[W]) const {
    if (W == 2) {
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(out[0]), "=d"(out[W - 1]) : "r"(a));
    } else {
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(out[0]) : "r"(a));
    }
  }
  template <int W>
  __device__ __forceinline__ void ld(int slot, const double *, long long i,
                                     double (&out)[W]) const {
    lds<W>(nb + off[slot] + (unsigned)(i - row0) * 8u, out);
  }
  template <int W>  // column j of the functor's column table
  __device__ __forceinline__ void ldc(int j, const double *, long long i,
                                      double (&out)[W]) const {
    lds<W>(nb + col0 + (unsigned)j * (ROWS * 8u) + (unsigned)(i - row0) * 8u, out);
  }
  template <int W>  // scratch slot of the stage (not filled by the copy engine)
  __device__ __forceinline__ void st(int slot, long long i, const double (&v)[W]) const {
    const unsigned a = nb + off[slot] + (unsigned)(i - row0) * 8u;
    if (W == 2) {
      asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v[0]), "d"(v[W - 1]) : "memory");
    } else {
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v[0]) : "memory");
    }
  }
  __device__ __forceinline__ double ldw(int slot, const double *, long long ci) const {
    double v[1];
    lds<1>(wb + (unsigned)(slot * wpitch) + (unsigned)(ci - con0) * 8u, v);
    return v[0];
  }
};

struct NoStreams {
  static constexpr int SRC = 0;   // 1: phases take a source (GSrc / SSrc) and slot ids
  static constexpr int TROWS = 128;  // rows per staged tile (few streams -> larger tiles)
  static constexpr int NB2 = 0;   // second-round per-constraint block sums (C2 / E phases)
  static constexpr int HASP = 0;  // P(ci, con): per-constraint prologue seen by A (AP form)
  static constexpr int NF = 0;    // third round F / FG after E: E leaves con.d[FD] to broadcast
  static constexpr int FD = 0;
  static constexpr int SMEM = 0;  // dynamic shared memory per block (bytes)
  // Staged kernels only: walk the tiles from the last one down.  Consecutive passes of an
  // iteration alternate their direction, so that a pass starts on the rows the pass
  // before it touched last -- the ~100 MB of them still in the 126 MB L2 (a few per cent
  // of a pass at 64M rows per GPU, a quarter of the small passes at 8M rows per GPU).
  static constexpr int REVERSE = 0;
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
template <class F, class S, class AT>
__device__ __forceinline__ void generic_range(const F &f, const S &src, const WDesc &w,
                                              long long lo, long long hi,
                                              long long tid, long long nthreads,
                                              AT &acc) {
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
      typename F::Con con;
      con.zero();
      if constexpr (F::SRC) {
        if constexpr (F::HASP) f.template AP<1>(src, i, coef, e, part, &acc, con);
        else f.template A<1>(src, i, coef, e, part, &acc);
      } else {
        if constexpr (F::HASP) f.template AP<1>(i, coef, e, part, &acc, con);
        else f.template A<1>(i, coef, e, part, &acc);
      }
      if constexpr (F::NB2 > 0) {
        double part2[1][F::NB2];
        if constexpr (F::SRC) f.template C2<1>(src, i, coef, e, con, acc, part2);
        else f.template C2<1>(i, coef, e, con, acc, part2);
      } else {
        if constexpr (F::SRC) f.template C<1>(src, i, coef, e, con, acc);
        else f.template C<1>(i, coef, e, con, acc);
      }
      if constexpr (F::NF > 0) {
        if constexpr (F::SRC) f.FG(src, i, coef[0], con, acc);
        else f.FG(i, coef[0], con, acc);
      }
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
      if constexpr (F::HASP) {
        if constexpr (F::SRC) f.P(src, ci, con);
        else f.P(ci, con);
      }
      for (int k = 0; k < w.nw; k++) {
        typename F::Elem e[1];
        double coef[1] = {k == 0 ? w.coef0 : w.coef_rest};
        double part[1][NB];
        if constexpr (F::SRC) {
          if constexpr (F::HASP) f.template AP<1>(src, j0 + k, coef, e, part, (AT *)nullptr, con);
          else f.template A<1>(src, j0 + k, coef, e, part, (AT *)nullptr);
        } else {
          if constexpr (F::HASP) f.template AP<1>(j0 + k, coef, e, part, (AT *)nullptr, con);
          else f.template A<1>(j0 + k, coef, e, part, (AT *)nullptr);
        }
#pragma unroll
        for (int b = 0; b < NB; b++) sum[b] += part[0][b];
      }
      if constexpr (F::SRC) f.B(src, ci, sum, con, acc);
      else f.B(ci, sum, con, acc);
      double sum2[F::NB2 > 0 ? F::NB2 : 1];
#pragma unroll
      for (int b = 0; b < (F::NB2 > 0 ? F::NB2 : 1); b++) sum2[b] = 0.0;
      for (int k = 0; k < w.nw; k++) {
        typename F::Elem e[1];
        double coef[1] = {k == 0 ? w.coef0 : w.coef_rest};
        double part[1][NB];
        if constexpr (F::SRC) {
          if constexpr (F::HASP) f.template AP<1>(src, j0 + k, coef, e, part, &acc, con);
          else f.template A<1>(src, j0 + k, coef, e, part, &acc);
        } else {
          if constexpr (F::HASP) f.template AP<1>(j0 + k, coef, e, part, &acc, con);
          else f.template A<1>(j0 + k, coef, e, part, &acc);
        }
        if constexpr (F::NB2 > 0) {
          double part2[1][F::NB2];
          if constexpr (F::SRC) f.template C2<1>(src, j0 + k, coef, e, con, acc, part2);
          else f.template C2<1>(j0 + k, coef, e, con, acc, part2);
#pragma unroll
          for (int b = 0; b < F::NB2; b++) sum2[b] += part2[0][b];
        } else {
          if constexpr (F::SRC) f.template C<1>(src, j0 + k, coef, e, con, acc);
          else f.template C<1>(j0 + k, coef, e, con, acc);
        }
      }
      if constexpr (F::NB2 > 0) {
        if constexpr (F::SRC) f.E(src, ci, sum2, con, acc);
        else f.E(ci, sum2, con, acc);
      }
      if constexpr (F::NF > 0) {
        for (int k = 0; k < w.nw; k++) {
          if constexpr (F::SRC) f.FG(src, j0 + k, k == 0 ? w.coef0 : w.coef_rest, con, acc);
          else f.FG(j0 + k, k == 0 ? w.coef0 : w.coef_rest, con, acc);
        }
      }
    }
  }
}

// One element pair (i, i+1) per lane, a full warp at a time: the shuffle path of
// the aligned power-of-two weighting blocks (w.mode == 1) or no blocks (mode 0).
template <class F, class S, class AT>
__device__ __forceinline__ void tile_pair(const F &f, const S &src, const WDesc &w,
                                          const long long i, const long long ncon_elems,
                                          const int half, AT &acc) {
  constexpr int NB = F::NB > 0 ? F::NB : 1;
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
  if constexpr (F::SRC) {
    if constexpr (F::HASP) {
      if (in_con) f.P(src, (i >> w.nw_log2), con);
      f.template AP<2>(src, i, coef, e, part, &acc, con);
    } else {
      f.template A<2>(src, i, coef, e, part, &acc);
    }
  } else {
    if constexpr (F::HASP) {
      if (in_con) f.P((i >> w.nw_log2), con);
      f.template AP<2>(i, coef, e, part, &acc, con);
    } else {
      f.template A<2>(i, coef, e, part, &acc);
    }
  }
  const int lane = threadIdx.x & 31;
  const int lead = lane & ~(half - 1);
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
    if (in_con && lane == lead) {
      if constexpr (F::SRC) f.B(src, (i >> w.nw_log2), sum, con, acc);
      else f.B((i >> w.nw_log2), sum, con, acc);
    }
    if (F::Con::ND > 0) {
#pragma unroll
      for (int b = 0; b < (F::Con::ND > 0 ? F::Con::ND : 1); b++)
        con.d[b] = __shfl_sync(0xffffffffu, con.d[b], lead);
    }
  }
  if constexpr (F::NB2 > 0) {
    double part2[2][F::NB2];
    if constexpr (F::SRC) f.template C2<2>(src, i, coef, e, con, acc, part2);
    else f.template C2<2>(i, coef, e, con, acc, part2);
    if (w.mode == 1) {
      double sum2[F::NB2];
#pragma unroll
      for (int b = 0; b < F::NB2; b++)
        sum2[b] = in_con ? part2[0][b] + part2[1][b] : 0.0;
      for (int o = 1; o < half; o <<= 1) {
#pragma unroll
        for (int b = 0; b < F::NB2; b++) sum2[b] += shfl_xor_d(sum2[b], o);
      }
      if (in_con && lane == lead) {
        if constexpr (F::SRC) f.E(src, (i >> w.nw_log2), sum2, con, acc);
        else f.E((i >> w.nw_log2), sum2, con, acc);
      }
      if constexpr (F::NF > 0)
        con.d[F::FD] = __shfl_sync(0xffffffffu, con.d[F::FD], lead);
    }
    if constexpr (F::NF > 0) {
      if constexpr (F::SRC) f.template F<2>(src, i, coef, e, con, acc);
      else f.template F<2>(i, coef, e, con, acc);
    }
  } else {
    if constexpr (F::SRC) f.template C<2>(src, i, coef, e, con, acc);
    else f.template C<2>(i, coef, e, con, acc);
  }
}

template <class F>
__global__ void __launch_bounds__(PCU_TILE_THREADS, F::MINB)
    tile_kernel(const F f, const long long n, const WDesc w, const RedBuf rb) {
  typename F::AccT acc;
  acc.init();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const GSrc src;

  if (w.mode == 2) {
    generic_range(f, src, w, 0, n, tid, nthreads, acc);
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
      tile_pair(f, src, w, i, ncon_elems, half, acc);
    }
    const long long tail_lo = 2 * nvec_main;
    if (tail_lo < n && blockIdx.x == 0) {
      WDesc wt = w;
      if (w.mode == 0) wt.nwcon = 0;
      generic_range(f, src, wt, tail_lo, n, (long long)threadIdx.x, (long long)blockDim.x, acc);
    }
  }
  if (F::NS + F::NX + F::NM > 0) {
    f.finalize(acc);
    finish_reduction<F::NS, F::NX, F::NM>(acc, rb);
  }
}

// ------------------------------------------------- bulk-copy staged harness
// The register-fed tile_kernel keeps (threads x loads in flight) bytes moving,
// and the fused KKT passes (35-45 streams, 70-100 registers) sit at 5.2-5.7 TB/s
// where a plain many-stream copy reaches 7 TB/s (profiles/r1b_stream_probe.txt).
// Here the copy engine does the streaming: NPW producer warps issue one
// cp.async.bulk per stream per tile (ROWS rows; the copies are dealt over the
// producer warps because one warp instruction serialises its lanes' copies)
// into a ring of shared-memory stages; consumer group g = (ROWS / 64 warps)
// owns the stages s with s % G == g and runs the functor's phases on them from
// shared memory (SSrc), storing results straight to global memory.  Bytes in
// flight = the ring, independent of registers; one CTA per SM, tiles dealt
// round robin.  Stages are complete/empty mbarrier pairs.
//   * W-streams (per weighting constraint) are staged for tiles that lie
//     entirely inside the blocks; the (at most one) straddling tile and the
//     ragged tail run through the global path on block 0 / the last block.
struct TmaPlan {
  long long ntiles;     // full tiles of ROWS rows
  long long tiles_con;  // tiles [0, tiles_con) lie inside the weighting blocks
  long long tile_skip;  // the straddling tile, or -1
  int groups;           // consumer groups (each ROWS / 64 warps)
  int nstages;          // multiple of groups, <= 16
  int stage_bytes;
  int woff;             // byte offset of the W-slots inside a stage
  int wpitch;           // bytes per W-slot
  int npw;              // producer warps
  int col_base;         // compact slot of the first column
  int reverse;          // 1: the tiles are taken from the last one down (F::REVERSE)
  unsigned long long nmap[3];  // fixed slot id -> compact slot, one byte each
  unsigned noff[24];           // the same as byte offsets inside a stage
};

#define PCU_TMA_NPW 4        // producer warps = warpgroup 0 (the first plan.npw of them issue copies)
#define PCU_TMA_PROD_REGS 24   // setmaxnreg: producers keep 24 registers ...
#define PCU_TMA_CONS_REGS 160  // ... the 12 consumer warps get 160 (launch: 128 each)
#define PCU_TMA_MAXWARPS 16  // 4 producer + 12 consumer warps per CTA (one CTA per SM)
#define PCU_TMA_MAXSTAGES 16
#define PCU_TMA_WPT 2         // consumer warps per group

__device__ __forceinline__ unsigned tt_smem_u32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tt_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tt_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tt_mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Waits for the phase with the given parity; the suspend-time hint lets the
// hardware park the warp instead of spinning through the issue slots the
// consumer warps on the same scheduler need (ncu: 13 % of ResF's executed
// instructions were try_wait / branch pairs without it).
__device__ __forceinline__ void tt_mbar_wait(unsigned bar, unsigned parity,
                                             unsigned hint_ns = 2000u) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TT_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra TT_DONE_%=;\n"
      "bra TT_WAIT_%=;\n"
      "TT_DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(hint_ns)
      : "memory");
}
__device__ __forceinline__ void tt_bulk_g2s(unsigned dst, const void *src,
                                            unsigned bytes, unsigned bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], "
      "%2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// Deals the functor's streams (tstreams order) over the producer threads:
// copy c goes to producer warp c % NPW, lane (c / NPW) % 32, at most two per thread.
template <int ROWS, int NFIX>
struct TmaAssign {
  unsigned long long nmap[3];
  int col_base;
  int me_warp, me_lane, c, npw;
  int cnt;
  const double *p0, *p1;
  unsigned off0, off1;  // byte offset inside the stage
  int w0, w1;           // 1: W-stream
  int woff, wpitch;
  __host__ __device__ __forceinline__ void take(const double *ptr, unsigned off, int isw) {
    if (ptr == nullptr) return;
    const int cc = c++;
    if (cc % npw != me_warp || (cc / npw) % 32 != me_lane) return;
    if (cnt == 0) {
      p0 = ptr; off0 = off; w0 = isw;
    } else {
      p1 = ptr; off1 = off; w1 = isw;
    }
    cnt++;
  }
  __host__ __device__ __forceinline__ void n(int slot, const double *ptr) {
    const int cs = slot < NFIX ? (int)((nmap[(slot >> 3) % 3] >> ((slot & 7) * 8)) & 0xffull)
                               : col_base + (slot - NFIX);
    take(ptr, (unsigned)cs * (ROWS * 8), 0);
  }
  __host__ __device__ __forceinline__ void w(int slot, const double *ptr) {
    take(ptr, (unsigned)(woff + slot * wpitch), 1);
  }
};

template <class F, int ROWS>
__global__ void __launch_bounds__(PCU_TMA_MAXWARPS * 32, 1)
    tma_tile_kernel(const F f, const long long n, const WDesc w, const RedBuf rb,
                    const TmaPlan plan) {
  constexpr int WPT = PCU_TMA_WPT;           // consumer warps per group
  constexpr int CHUNKS = ROWS / (64 * WPT);  // 64-row chunks per warp per tile
  constexpr int NCWMAX = PCU_TMA_MAXWARPS - PCU_TMA_NPW;
  extern __shared__ double2 pcu_dyn_smem[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(pcu_dyn_smem);
  __shared__ __align__(8) unsigned long long tt_full[PCU_TMA_MAXSTAGES];
  __shared__ __align__(8) unsigned long long tt_empty[PCU_TMA_MAXSTAGES];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = plan.nstages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) {
      tt_mbar_init(tt_smem_u32(&tt_full[s]), plan.npw);
      tt_mbar_init(tt_smem_u32(&tt_empty[s]), WPT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const long long ncon_elems = (w.mode == 1) ? (long long)w.nwcon * w.nw : 0;
  const int half = (w.mode == 1) ? (w.nw >> 1) : 1;
  const int con_per_tile = (w.mode == 1) ? ROWS / w.nw : 0;
  const unsigned full0 = tt_smem_u32(tt_full), empty0 = tt_smem_u32(tt_empty);
  const unsigned smem0 = tt_smem_u32(smem);
  // this CTA's tiles are blockIdx.x + j gridDim.x, j < ntl_all; the straddling tile
  // (position js in that sequence, or -1) is left to the global path
  const int ntl_all =
      (long long)blockIdx.x < plan.ntiles ? (int)((plan.ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  // position q = blockIdx.x + j gridDim.x of the sequence is tile q, or ntiles - 1 - q
  int js = -1;
  const long long skip_q = plan.tile_skip < 0 ? -1
                           : (plan.reverse ? plan.ntiles - 1 - plan.tile_skip : plan.tile_skip);
  if (skip_q >= (long long)blockIdx.x &&
      (skip_q - (long long)blockIdx.x) % (long long)gridDim.x == 0)
    js = (int)((skip_q - (long long)blockIdx.x) / (long long)gridDim.x);
  const int ntl = ntl_all - (js >= 0 ? 1 : 0);

  if (warp < PCU_TMA_NPW) {
    // ------------------------------------------------- producers (warpgroup 0)
    // hand most of this warpgroup's registers to the consumers
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PCU_TMA_PROD_REGS));
    if (warp >= plan.npw) return;
    TmaAssign<ROWS, F::NFIX> as;
    as.nmap[0] = plan.nmap[0];
    as.nmap[1] = plan.nmap[1];
    as.nmap[2] = plan.nmap[2];
    as.col_base = plan.col_base;
    as.me_warp = warp;
    as.me_lane = lane;
    as.npw = plan.npw;
    as.c = 0;
    as.cnt = 0;
    as.p0 = as.p1 = nullptr;
    as.off0 = as.off1 = 0;
    as.w0 = as.w1 = 0;
    as.woff = plan.woff;
    as.wpitch = plan.wpitch;
    f.tstreams(as);
    const unsigned nbytes = ROWS * 8, wbytes = (unsigned)con_per_tile * 8;
    unsigned tx_n = 0, tx_w = 0;
    if (as.cnt > 0) { if (as.w0) tx_w += wbytes; else tx_n += nbytes; }
    if (as.cnt > 1) { if (as.w1) tx_w += wbytes; else tx_n += nbytes; }
    for (int o = 16; o > 0; o >>= 1) {
      tx_n += __shfl_xor_sync(0xffffffffu, tx_n, o);
      tx_w += __shfl_xor_sync(0xffffffffu, tx_w, o);
    }
    int s = 0;
    unsigned round = 0;
    for (int kt = 0; kt < ntl; kt++) {
      const int j = (js >= 0 && kt >= js) ? kt + 1 : kt;
      const long long tq = blockIdx.x + (long long)j * gridDim.x;
      const long long tile = plan.reverse ? plan.ntiles - 1 - tq : tq;
      if (round > 0) tt_mbar_wait(empty0 + 8u * s, (round - 1) & 1);
      const bool in_con = tile < plan.tiles_con;
      const unsigned full = full0 + 8u * s;
      if (lane == 0) tt_mbar_expect_tx(full, tx_n + (in_con ? tx_w : 0u));
      __syncwarp();
      const unsigned base = smem0 + (unsigned)s * (unsigned)plan.stage_bytes;
      if (as.cnt > 0) {
        if (!as.w0) tt_bulk_g2s(base + as.off0, as.p0 + tile * ROWS, nbytes, full);
        else if (in_con) tt_bulk_g2s(base + as.off0, as.p0 + tile * con_per_tile, wbytes, full);
      }
      if (as.cnt > 1) {
        if (!as.w1) tt_bulk_g2s(base + as.off1, as.p1 + tile * ROWS, nbytes, full);
        else if (in_con) tt_bulk_g2s(base + as.off1, as.p1 + tile * con_per_tile, wbytes, full);
      }
      if (++s == S) {
        s = 0;
        round++;
      }
    }
    return;
  }
  // ------------------------------------------------- consumers (warpgroups 1..3)
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PCU_TMA_CONS_REGS));
  typedef Acc<F::NS, F::NX, F::NM> AT;
  AT acc;
  acc.init();
  const int cw = warp - PCU_TMA_NPW;   // consumer warp index
  const int ncw = plan.groups * WPT;   // consumer warps with a group
  if (cw < ncw) {
    // group g owns the tiles with sequence index g, g + G, ... and the stages
    // g + G d (d cycles through the group's ring depth)
    const int g = cw / WPT, wg = cw % WPT;
    const int G = plan.groups, depth = S / G;
    int d = 0;
    unsigned ph = 0;
    for (int kt = g; kt < ntl; kt += G) {
      const int j = (js >= 0 && kt >= js) ? kt + 1 : kt;
      const long long tq = blockIdx.x + (long long)j * gridDim.x;
      const long long tile = plan.reverse ? plan.ntiles - 1 - tq : tq;
      const int s = g + G * d;
      tt_mbar_wait(full0 + 8u * s, ph);
      SSrc<ROWS, F::NFIX> src;
      src.nb = smem0 + (unsigned)s * (unsigned)plan.stage_bytes;
      src.wb = src.nb + (unsigned)plan.woff;
      src.row0 = tile * ROWS;
      src.con0 = tile * con_per_tile;
      src.wpitch = plan.wpitch;
#pragma unroll
      for (int q = 0; q < F::NFIX; q++) src.off[q] = plan.noff[q];
      src.col0 = (unsigned)plan.col_base * (ROWS * 8u);
#pragma unroll 1
      for (int c = 0; c < CHUNKS; c++)
        tile_pair(f, src, w, src.row0 + (c * WPT + wg) * 64 + 2 * lane, ncon_elems, half, acc);
      __syncwarp();
      if (lane == 0) tt_mbar_arrive(empty0 + 8u * s);
      if (++d == depth) {
        d = 0;
        ph ^= 1u;
      }
    }
    // what the staged loop left out, through the global path
    const GSrc gsrc;
    if (plan.tile_skip >= 0 && blockIdx.x == gridDim.x - 1) {
      for (int v = cw * 32 + lane; v < ROWS / 2; v += ncw * 32)
        tile_pair(f, gsrc, w, plan.tile_skip * ROWS + 2 * v, ncon_elems, half, acc);
    }
    if (blockIdx.x == 0) {
      const long long nvec_main = ((n / 2) / 32) * 32;
      for (long long v = plan.ntiles * (ROWS / 2) + cw * 32 + lane; v < nvec_main;
           v += ncw * 32)
        tile_pair(f, gsrc, w, 2 * v, ncon_elems, half, acc);
      const long long tail_lo = 2 * nvec_main;
      if (tail_lo < n) {
        WDesc wt = w;
        if (w.mode == 0) wt.nwcon = 0;
        generic_range(f, gsrc, wt, tail_lo, n, (long long)(cw * 32 + lane),
                      (long long)(ncw * 32), acc);
      }
    }
  }
  if (F::NS + F::NX + F::NM > 0) {
    f.finalize(acc);
    finish_reduction<F::NS, F::NX, F::NM, AT, NCWMAX>(acc, rb, (int)threadIdx.x - PCU_TMA_NPW * 32,
                                                       NCWMAX * 32);
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
