// pcu_gram.cu -- weighted tall-skinny Gram  S = V^T P V  on the FP64 tensor cores.
//
// V = [A | Z] are the (c + q) column vectors (dense constraint gradients and the
// compact quasi-Newton vectors), P = D0^-1 restricted to the design variables:
//     P = Dinv - Dinv Aw^T Ew^-1 Aw Dinv          (SM.cpp:122-190, nwblock = 1)
// One pass over the columns replaces the reference's c + q sequential
// mat->apply / solveKKTDiagSystem calls and their c(c+1)/2 + q^2 separately
// reduced dot products (IP.cpp:1932-1961 and 2646-2661).
//
// Tensor-core mapping: mma.sync.aligned.m8n8k4 .f64 (DMMA; tcgen05 has no fp64
// kind).  A lane (gi = lane/4, kk = lane%4) loads, for every 8-column tile T, the
// two rows r0 + 2kk, r0 + 2kk + 1 of column 8T + gi with one 128-bit access.
// That register pair is simultaneously the B fragment (k = kk, n = gi) of tile
// T and, multiplied by the row weight, the A fragment (m = gi, k = kk): the
// reduction index may be permuted freely as long as A and B use the same
// permutation, so two DMMAs (.x rows, .y rows) consume 8 rows.
// The weighting-constraint correction  - sum_i Cw_i u_i u_i^T,
// u_i = sum_{r in block i} coef_r Dinv_r V[r,:], is formed in registers with
// two shuffles and issued as one more DMMA with a single non-zero k slot.
#include <stdlib.h>
#include <string.h>

#include <utility>
#include <vector>

#include "pcu_ctx.cuh"

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, "
      "{%0,%1};"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

template <int NTA, int NTB, bool DIAG>
struct GramPairs {
  static constexpr int N = DIAG ? (NTA * (NTA + 1)) / 2 : NTA * NTB;
};

// result layout: col-major mpad x mpad, tile (Ti, Tj) written at rows 8Ti..,
// cols 8Tj.. ; DIAG computes tiles with Ti >= Tj of block A.
template <int NTA, int NTB, bool DIAG>
__global__ void __launch_bounds__(PCU_THREADS)
    gram_kernel(const ColTable cols, const int colA0, const int colB0,
                const int m, const double *__restrict__ Dinv,
                const double *__restrict__ Cw, const WDesc w, const long long n,
                double *__restrict__ partials, unsigned int *counter,
                double *__restrict__ result, const int ld,
                const long long list0, const long long list1, const int nlist,
                const int accumulate, const double *__restrict__ d2,
                const int rhs_col) {
  // nlist > 0: process only the chunks list0 [, list1] and add to `result`
  // rhs_col >= 0: see gram_fast_kernel (the A-side block sum of that column is
  // shifted by -d2)
  constexpr int NP = GramPairs<NTA, NTB, DIAG>::N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gi = lane >> 2, kk = lane & 3;
  const int nwarps_cta = blockDim.x >> 5;
  const long long gwarp = (long long)blockIdx.x * nwarps_cta + warp;
  const long long nwarps = (long long)gridDim.x * nwarps_cta;

  double acc[NP][2];
#pragma unroll
  for (int p = 0; p < NP; p++) acc[p][0] = acc[p][1] = 0.0;
  double uA[NTA], uB[NTB], hA[NTA], hB[NTB], hcw = 0.0;
  bool rhsA[NTA];
#pragma unroll
  for (int t = 0; t < NTA; t++) rhsA[t] = (rhs_col >= 0) && (colA0 + 8 * t + gi == rhs_col);
#pragma unroll
  for (int t = 0; t < NTA; t++) uA[t] = hA[t] = 0.0;
#pragma unroll
  for (int t = 0; t < NTB; t++) uB[t] = hB[t] = 0.0;

  const bool wmode = (w.mode == 1);
  const long long ncon_elems = wmode ? (long long)w.nwcon * w.nw : 0;
  const int L = wmode ? ((w.nw < 8 ? w.nw : 8) >> 1) : 1;  // lanes per block

  const double *colA[NTA];
  const double *colB[NTB];
#pragma unroll
  for (int t = 0; t < NTA; t++) {
    const int c = colA0 + 8 * t + gi;
    colA[t] = (c < m) ? cols.p[c] : nullptr;
  }
#pragma unroll
  for (int t = 0; t < NTB; t++) {
    const int c = colB0 + 8 * t + gi;
    colB[t] = (c < m) ? cols.p[c] : nullptr;
  }

  // Software-pipelined walk over this warp's 64-row chunks, 8 rows per step:
  // the loads of step s+1 are issued before the DMMAs of step s.
  const long long nchunks = nlist > 0 ? nlist : (n + 63) / 64;
  auto chunk_row = [&](long long c) -> long long {
    return 64 * (nlist > 0 ? (c == 0 ? list0 : list1) : c);
  };
  struct Frag {
    double2 wv;
    double2 fb[NTB];
    double2 fa[DIAG ? 1 : NTA];
  };
  auto load_step = [&](long long r0, Frag &f) {
    const long long r = r0 + 2 * kk;
    const bool ok0 = r < n, ok1 = (r + 1) < n;
    f.wv = make_double2(0.0, 0.0);
    if (ok1) {
      f.wv = Dinv ? *reinterpret_cast<const double2 *>(Dinv + r)
                  : make_double2(1.0, 1.0);
    } else if (ok0) {
      f.wv.x = Dinv ? Dinv[r] : 1.0;
    }
#pragma unroll
    for (int t = 0; t < NTB; t++) {
      double2 v = make_double2(0.0, 0.0);
      if (colB[t]) {
        if (ok1) v = *reinterpret_cast<const double2 *>(colB[t] + r);
        else if (ok0) v.x = colB[t][r];
      }
      f.fb[t] = v;
    }
    if (!DIAG) {
#pragma unroll
      for (int t = 0; t < NTA; t++) {
        double2 v = make_double2(0.0, 0.0);
        if (colA[t]) {
          if (ok1) v = *reinterpret_cast<const double2 *>(colA[t] + r);
          else if (ok0) v.x = colA[t][r];
        }
        f.fa[t] = v;
      }
    }
  };

  long long chunk = gwarp;
  int step = 0;
  bool have = chunk < nchunks;
  Frag cur, nxt;
  if (have) load_step(chunk_row(chunk), cur);
  while (have) {
    long long nchunk = chunk;
    int nstep = step + 1;
    if (nstep == 8) {
      nstep = 0;
      nchunk += nwarps;
    }
    const long long r0 = chunk_row(chunk) + step * 8;
    bool nhave = nchunk < nchunks;
    if (nhave && chunk_row(nchunk) + nstep * 8 >= n) {  // ragged last chunk
      if (nlist > 0 && nstep != 0) {
        nstep = 0;  // rest of this listed chunk is past the end: go to the next one
        nchunk += nwarps;
        nhave = nchunk < nchunks && chunk_row(nchunk) < n;
      } else {
        nhave = false;
      }
    }
    if (nhave) load_step(chunk_row(nchunk) + nstep * 8, nxt);
    {
      const long long r = r0 + 2 * kk;
      const double2 wv = cur.wv;
      double2 fa[NTA];
#pragma unroll
      for (int t = 0; t < NTA; t++) {
        const double2 f = DIAG ? cur.fb[t < NTB ? t : 0] : cur.fa[DIAG ? 0 : t];
        fa[t] = make_double2(f.x * wv.x, f.y * wv.y);
      }
      const double2 *fb = cur.fb;
      {
        int p = 0;
#pragma unroll
        for (int ti = 0; ti < NTA; ti++) {
#pragma unroll
          for (int tj = 0; tj < NTB; tj++) {
            if (!DIAG || tj <= ti) {
              dmma884(acc[p], fa[ti].x, fb[tj].x);
              p++;
            }
          }
        }
        // second half of the 8 rows: the two DMMAs on one accumulator are kept
        // NP instructions apart so they never issue back to back
        p = 0;
#pragma unroll
        for (int ti = 0; ti < NTA; ti++) {
#pragma unroll
          for (int tj = 0; tj < NTB; tj++) {
            if (!DIAG || tj <= ti) {
              dmma884(acc[p], fa[ti].y, fb[tj].y);
              p++;
            }
          }
        }
      }
      if (wmode) {
        const bool in_con = r < ncon_elems;
        const double c0 = ((r & (long long)(w.nw - 1)) == 0) ? w.coef0 : w.coef_rest;
        const double c1 = w.coef_rest;
#pragma unroll
        for (int t = 0; t < NTA; t++)
          uA[t] += in_con ? fma(fa[t].x, c0, fa[t].y * c1) : 0.0;
        if (!DIAG) {
#pragma unroll
          for (int t = 0; t < NTB; t++)
            uB[t] += in_con ? fma(fb[t].x * wv.x, c0, fb[t].y * wv.y * c1) : 0.0;
        }
        const bool block_end = (w.nw <= 8) || (((r0 + 8) & (long long)(w.nw - 1)) == 0);
        if (block_end) {
          for (int o = 1; o < L; o <<= 1) {
#pragma unroll
            for (int t = 0; t < NTA; t++) uA[t] += shfl_xor_d(uA[t], o);
            if (!DIAG) {
#pragma unroll
              for (int t = 0; t < NTB; t++) uB[t] += shfl_xor_d(uB[t], o);
            }
          }
          if (w.nw >= 8) {
            // one block per >= 1 steps: park its u in k-slot (block & 3) and issue
            // ONE correction DMMA per four blocks (all four k slots in use)
            const int blk = (int)((r0 / w.nw) & 3);
            if (kk == blk) {
              hcw = in_con ? Cw[r0 / w.nw] : 0.0;
#pragma unroll
              for (int t = 0; t < NTA; t++)
                hA[t] = in_con ? (rhsA[t] ? uA[t] - d2[r0 / w.nw] : uA[t]) : 0.0;
#pragma unroll
              for (int t = 0; t < NTB; t++)
                hB[t] = in_con ? (DIAG ? uA[t < NTA ? t : 0] : uB[t]) : 0.0;
            }
            const bool flush = (blk == 3) || (step == 7) || (r0 + 8 >= n);
            if (flush) {
              int p = 0;
#pragma unroll
              for (int ti = 0; ti < NTA; ti++) {
#pragma unroll
                for (int tj = 0; tj < NTB; tj++) {
                  if (!DIAG || tj <= ti) {
                    dmma884(acc[p], -hcw * hA[ti], hB[tj]);
                    p++;
                  }
                }
              }
              hcw = 0.0;
#pragma unroll
              for (int t = 0; t < NTA; t++) hA[t] = 0.0;
#pragma unroll
              for (int t = 0; t < NTB; t++) hB[t] = 0.0;
            }
          } else {
            const bool lead = in_con && ((kk & (L - 1)) == 0);
            const double cw = lead ? Cw[r / w.nw] : 0.0;
            double ua[NTA], ub[NTB];
#pragma unroll
            for (int t = 0; t < NTA; t++)
              ua[t] = lead ? -cw * (rhsA[t] ? uA[t] - d2[r / w.nw] : uA[t]) : 0.0;
#pragma unroll
            for (int t = 0; t < NTB; t++)
              ub[t] = lead ? (DIAG ? uA[t < NTA ? t : 0] : uB[t]) : 0.0;
            int p = 0;
#pragma unroll
            for (int ti = 0; ti < NTA; ti++) {
#pragma unroll
              for (int tj = 0; tj < NTB; tj++) {
                if (!DIAG || tj <= ti) {
                  dmma884(acc[p], ua[ti], ub[tj]);
                  p++;
                }
              }
            }
          }
#pragma unroll
          for (int t = 0; t < NTA; t++) uA[t] = 0.0;
#pragma unroll
          for (int t = 0; t < NTB; t++) uB[t] = 0.0;
        }
      }
    }
    cur = nxt;
    chunk = nchunk;
    step = nstep;
    have = nhave;
  }

  // ---- CTA combine (pair by pair), then grid combine by the last block ----
  __shared__ double sm[PCU_THREADS / 32][64];
  __shared__ bool is_last;
#pragma unroll
  for (int p = 0; p < NP; p++) {
    // element (row gi, col 2kk+e) of the 8x8 tile, stored col-major in sm
    sm[warp][gi + 8 * (2 * kk)] = acc[p][0];
    sm[warp][gi + 8 * (2 * kk + 1)] = acc[p][1];
    __syncthreads();
    if (threadIdx.x < 64) {
      double v = 0.0;
      for (int ww = 0; ww < nwarps_cta; ww++) v += sm[ww][threadIdx.x];
      partials[((size_t)blockIdx.x * NP + p) * 64 + threadIdx.x] = v;
    }
    __syncthreads();
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
    for (int idx = threadIdx.x; idx < NP * 64; idx += blockDim.x) {
      const double v = pcu_ordered_sum(partials + idx, (size_t)NP * 64, 0u, 1u, gridDim.x);
      // decode pair -> (ti, tj)
      const int p = idx >> 6, e = idx & 63;
      int ti = 0, tj = 0;
      if (DIAG) {
        int q = p;
        while (q > ti) {
          q -= ti + 1;
          ti++;
        }
        tj = q;
      } else {
        ti = p / NTB;
        tj = p % NTB;
      }
      const int row = colA0 + 8 * ti + (e & 7);
      const int col = colB0 + 8 * tj + (e >> 3);
      if (row < ld && col < ld) {
        if (accumulate) result[(size_t)row + (size_t)ld * col] += v;
        else result[(size_t)row + (size_t)ld * col] = v;
      }
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

#include "pcu_gram_fast.cuh"
#include "pcu_gram_tma.cuh"

// Compatibility path for weighting patterns the shuffle layout cannot express
// (WDesc.mode == 2, e.g. the 5-of-6 pattern of examples/rosenbrock): subtracts
// sum_i Cw_i u_i u_i^T from the lower triangle.  One thread per (row, col) entry;
// meant for small W.
__global__ void gram_generic_correction(const ColTable cols, int m,
                                        const double *__restrict__ Dinv,
                                        const double *__restrict__ Cw,
                                        const WDesc w, double *result, int ld,
                                        const double *__restrict__ d2,
                                        int rhs_col) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * m) return;
  const int i = idx % m, j = idx / m;
  if (j > i) return;
  double s = 0.0;
  for (int ci = 0; ci < w.nwcon; ci++) {
    const long long j0 = w.wstart + (long long)ci * w.wstride;
    double ui = 0.0, uj = 0.0;
    for (int k = 0; k < w.nw; k++) {
      const double cf = (k == 0 ? w.coef0 : w.coef_rest);
      const double d = Dinv ? Dinv[j0 + k] : 1.0;
      ui = fma(cf * d, cols.p[i][j0 + k], ui);
      uj = fma(cf * d, cols.p[j][j0 + k], uj);
    }
    if (i == rhs_col) ui -= d2[ci];
    s = fma(Cw[ci] * ui, uj, s);
  }
  result[(size_t)i + (size_t)ld * j] -= s;
}

template <int NTA, int NTB, bool DIAG>
static int launch_gram(pcu_ctx *ctx, const ColTable &cols, int colA0, int colB0,
                       int m, const double *Dinv, const double *Cw,
                       const WDesc &w, long long n, double *result, int ld,
                       long long list0 = 0, long long list1 = 0, int nlist = 0,
                       int accumulate = 0, const double *d2 = nullptr,
                       int rhs_col = -1) {
  constexpr int NP = GramPairs<NTA, NTB, DIAG>::N;
  long long nchunks = (n + 63) / 64;
  long long need = (nchunks + (PCU_THREADS / 32) - 1) / (PCU_THREADS / 32);
  static int blocks_per_sm = -1;
  if (blocks_per_sm < 0) {
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &v, gram_kernel<NTA, NTB, DIAG>, PCU_THREADS, 0) != cudaSuccess || v < 1)
      v = 1;
    blocks_per_sm = v > 4 ? 4 : v;
  }
  int grid = ctx->num_sms * blocks_per_sm;
  if (need < grid) grid = (int)(need < 1 ? 1 : need);
  if (nlist > 0) grid = 1;
  if (ctx->big_reserve(0, (size_t)grid * NP * 64)) return 1;
  ctx->prof_begin("gram_kernel");
  gram_kernel<NTA, NTB, DIAG><<<grid, PCU_THREADS, 0, ctx->stream>>>(
      cols, colA0, colB0, m, Dinv, Cw, w, n, ctx->d_big_partials, ctx->d_counter,
      result, ld, list0, list1, nlist, accumulate, d2, rhs_col);
  ctx->prof_end();
  ctx->launches++;
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int NT, int NWC>
static int launch_gram_fast_t(pcu_ctx *ctx, const ColTable &cols, int m,
                              const double *Dinv, const double *Cw,
                              const WDesc &w, long long n, double *result,
                              int ld, const double *d2, int rhs_col) {
  constexpr int NP = (NT * (NT + 1)) / 2;
  static int blocks_per_sm = -1;
  if (blocks_per_sm < 0) {
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &v, gram_fast_kernel<NT, NWC>, PCU_THREADS, 0) != cudaSuccess || v < 1)
      v = 1;
    blocks_per_sm = v > 4 ? 4 : v;
  }
  const long long nchunks = n / 64;
  long long need = (nchunks + (PCU_THREADS / 32) - 1) / (PCU_THREADS / 32);
  int grid = ctx->num_sms * blocks_per_sm;
  if (need < grid) grid = (int)(need < 1 ? 1 : need);
  if (ctx->big_reserve(0, (size_t)grid * NP * 64)) return 1;
  ctx->prof_begin("gram_kernel");
  gram_fast_kernel<NT, NWC><<<grid, PCU_THREADS, 0, ctx->stream>>>(
      cols, m, Dinv, Cw, w, n, ctx->d_big_partials, ctx->d_counter, result, ld, d2,
      rhs_col);
  ctx->prof_end();
  ctx->launches++;
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int NT>
static int launch_gram_fast_n(pcu_ctx *ctx, const ColTable &cols, int m,
                              const double *Dinv, const double *Cw,
                              const WDesc &w, long long n, double *result,
                              int ld, int nwc, const double *d2, int rhs_col) {
  if (nwc == 0) return launch_gram_fast_t<NT, 0>(ctx, cols, m, Dinv, Cw, w, n, result, ld, d2, rhs_col);
  if (nwc == 8) return launch_gram_fast_t<NT, 8>(ctx, cols, m, Dinv, Cw, w, n, result, ld, d2, rhs_col);
  return launch_gram_fast_t<NT, -1>(ctx, cols, m, Dinv, Cw, w, n, result, ld, d2, rhs_col);
}

static int launch_gram_fast(pcu_ctx *ctx, const ColTable &cols, int m,
                            const double *Dinv, const double *Cw, const WDesc &w,
                            long long n, double *result, int ld, int nt, int nwc,
                            const double *d2, int rhs_col) {
  switch (nt) {
    case 1: return launch_gram_fast_n<1>(ctx, cols, m, Dinv, Cw, w, n, result, ld, nwc, d2, rhs_col);
    case 2: return launch_gram_fast_n<2>(ctx, cols, m, Dinv, Cw, w, n, result, ld, nwc, d2, rhs_col);
    case 3: return launch_gram_fast_n<3>(ctx, cols, m, Dinv, Cw, w, n, result, ld, nwc, d2, rhs_col);
    case 4: return launch_gram_fast_n<4>(ctx, cols, m, Dinv, Cw, w, n, result, ld, nwc, d2, rhs_col);
    default: return launch_gram_fast_n<5>(ctx, cols, m, Dinv, Cw, w, n, result, ld, nwc, d2, rhs_col);
  }
}

template <int NT, int NWC, int NCW, int RPW>
static int launch_gram_tma_t(pcu_ctx *ctx, const ColTable &cols, int m,
                             const double *Dinv, const double *Cw, const WDesc &w,
                             long long n, double *result, int ld, const double *d2,
                             int rhs_col, long long *rows_done, long long *skip_lo,
                             int *slab_rows) {
  constexpr int NP = (NT * (NT + 1)) / 2;
  constexpr int ROWS = RPW * NCW;
  *slab_rows = ROWS;
  int stage_bytes = (m + 1) * (ROWS * 8 + 64) + 2 * ROWS;
  stage_bytes = (stage_bytes + 127) / 128 * 128;
  // 227 KB per block less the kernel's 8.3 KB of static shared memory
  int nstages = (214 * 1024) / stage_bytes;
  if (nstages > PCU_GT_MAXSTAGES) nstages = PCU_GT_MAXSTAGES;
  if (nstages < 2) return -1;
  if (const char *e = getenv("PCU_GT_STAGES")) {
    const int v = atoi(e);
    if (v >= 2 && v <= nstages) nstages = v;
  }
  // whole slabs; the one that straddles the end of the weighting blocks is skipped
  const long long nslabs = n / ROWS;
  long long slab_con = 0, slab_skip = -1;
  if (w.mode == 1) {
    const long long ncon_elems = (long long)w.nwcon * w.nw;
    slab_con = ncon_elems / ROWS;
    if (slab_con > nslabs) slab_con = nslabs;
    if (ncon_elems % ROWS != 0 && slab_con < nslabs) slab_skip = slab_con;
  }
  *rows_done = nslabs * ROWS;
  *skip_lo = slab_skip >= 0 ? slab_skip * ROWS : -1;
  const int smem = nstages * stage_bytes;
  static int attr_smem_dev[PCU_MAX_DEVICES] = {0};  // per device (function attribute)
  int &attr_smem = attr_smem_dev[ctx->device >= 0 && ctx->device < PCU_MAX_DEVICES ? ctx->device : 0];
  if (smem > attr_smem || ctx->device >= PCU_MAX_DEVICES) {
    PCU_CUDA_OK(cudaFuncSetAttribute(gram_tma_kernel<NT, NWC, NCW, RPW>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  int grid = ctx->num_sms;
  if (nslabs < grid) grid = (int)nslabs;
  if (ctx->big_reserve(0, (size_t)grid * NP * 64)) return 1;
  ctx->prof_begin("gram_kernel");
  gram_tma_kernel<NT, NWC, NCW, RPW><<<grid, PCU_GT_THREADS_T(NCW), smem, ctx->stream>>>(
      cols, m, Dinv, Cw, w, nslabs, slab_con, slab_skip, nstages, stage_bytes,
      ctx->d_big_partials, ctx->d_counter, result, ld, d2, rhs_col, ctx->no_reverse ? 0 : 1);
  ctx->prof_end();
  ctx->launches++;
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

static int launch_gram_tma(pcu_ctx *ctx, const ColTable &cols, int m,
                           const double *Dinv, const double *Cw, const WDesc &w,
                           long long n, double *result, int ld, const double *d2,
                           int rhs_col, int nt, int nwc, long long *rows_done,
                           long long *skip_lo, int *slab_rows) {
#define PCU_GT_ARGS \
  ctx, cols, m, Dinv, Cw, w, n, result, ld, d2, rhs_col, rows_done, skip_lo, slab_rows
#define PCU_GT_CASE(NT_, NCW_)                                              \
  case NT_:                                                                 \
    return nwc == 0 ? launch_gram_tma_t<NT_, 0, NCW_, 32>(PCU_GT_ARGS)      \
                    : launch_gram_tma_t<NT_, 8, NCW_, 32>(PCU_GT_ARGS);
  // 17-24 columns with 16 consumer warps: 24 rows per warp make a stage 73 KB instead
  // of 97 KB, and three of them fit (PCU_GT_RPW=32 restores the two-stage ring)
  static const int rpw3 = getenv("PCU_GT_RPW") ? atoi(getenv("PCU_GT_RPW")) : PCU_GT_RPW3;
  if (nt == 3 && rpw3 == 24) {
    return nwc == 0 ? launch_gram_tma_t<3, 0, 16, 24>(PCU_GT_ARGS)
                    : launch_gram_tma_t<3, 8, 16, 24>(PCU_GT_ARGS);
  }
  switch (nt) {
    PCU_GT_CASE(1, 16)
    PCU_GT_CASE(2, 16)
    PCU_GT_CASE(3, 16)
    PCU_GT_CASE(4, 8)
    PCU_GT_CASE(5, 8)
  }
  return -1;
#undef PCU_GT_CASE
#undef PCU_GT_ARGS
}

// Wide path: one launch for up to 160 columns, no weighting correction.
template <int N2U, int SIDE>
static int launch_gram_wide_t(pcu_ctx *ctx, const ColTable &cols, int m, int nt,
                              const GramSegTable &segs, const double *Dinv, long long nslabs,
                              int nstages, int stage_bytes, double *result, int ld,
                              int side_off) {
  const int smem = nstages * stage_bytes;
  static int attr_smem_dev[PCU_MAX_DEVICES] = {0};  // per device (function attribute)
  int &attr_smem = attr_smem_dev[ctx->device >= 0 && ctx->device < PCU_MAX_DEVICES ? ctx->device : 0];
  if (smem > attr_smem || ctx->device >= PCU_MAX_DEVICES) {
    PCU_CUDA_OK(cudaFuncSetAttribute(gram_wide_kernel<N2U, SIDE>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  int grid = ctx->num_sms;
  if (nslabs < grid) grid = (int)nslabs;
  const int npairs = nt * (nt + 1) / 2;
  if (ctx->big_reserve(0, (size_t)grid * ((size_t)npairs * 64 + (SIDE ? ld : 0)))) return 1;
  ctx->prof_begin("gram_kernel");
  gram_wide_kernel<N2U, SIDE><<<grid, 32 * (PCU_GW_NCW + PCU_GW_NPW), smem, ctx->stream>>>(
      cols, m, nt, segs, Dinv, nslabs, nstages, stage_bytes, ctx->d_big_partials,
      ctx->d_counter, result, ld, ctx->no_reverse ? 0 : 1, side_off);
  ctx->prof_end();
  ctx->launches++;
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

// Work partition of the wide kernel for m columns (host only): tile rows nt, common
// two-pair segments per warp n2u, whether the last column is a side column, and the
// segment table.  Returns -1 when the kernel does not take this width.
static int gram_wide_deal(int m, int *nt_out, int *n2u_out, bool *side_out, GramSegTable *segs_out) {
  // a last column alone in its tile (m = 8 k + 1: C4's 120 columns + right-hand side) is
  // taken as the side column of two tile rows per warp instead of a tile row of its own
  static const bool no_side = getenv("PCU_NO_GRAM_SIDE") != nullptr;
  const bool side = !no_side && m > 8 && (m % 8) == 1;
  const int nt = side ? (m - 1) / 8 : (m + 7) / 8;
  if (side && nt > 2 * PCU_GW_NCW) return -1;  // warp w takes tile rows w and w + NCW
  // segments of the tile triangle: two-pair (ti, tj0), (ti, tj0 + 1) and, at the end of
  // the odd rows, single-pair ones
  std::vector<std::pair<int, int> > two, one;
  for (int ti = nt - 1; ti >= 0; ti--)
    for (int tj = 0; tj <= ti; tj += 2) (tj + 1 <= ti ? two : one).push_back({ti, tj});
  const int n2u = (int)two.size() / PCU_GW_NCW;  // common two-pair segments per warp
  if (n2u > PCU_GW_MAXN2U) return -1;
  GramSegTable &segs = *segs_out;
  memset(&segs, 0, sizeof(segs));
  int load[PCU_GW_NCW] = {0};
  size_t at = 0;
  auto put = [&](int w, int slot, const std::pair<int, int> &sg, int np) {
    segs.ti[w][slot] = (unsigned char)sg.first;
    segs.tj[w][slot] = (unsigned char)sg.second;
    segs.np[w][slot] = (unsigned char)np;
    load[w] += np;
  };
  for (int w = 0; w < PCU_GW_NCW; w++)
    for (int s = 0; s < n2u; s++, at++) put(w, s, two[at], 2);
  // the remaining two-pair segments, then the single-pair ones: each to the warp with the
  // fewest pairs that still has one of its two optional slots (n2u, n2u + 1) free
  auto place = [&](const std::pair<int, int> &sg, int np) -> bool {
    int w = -1;
    for (int c = 0; c < PCU_GW_NCW; c++)
      if ((segs.np[c][n2u] == 0 || segs.np[c][n2u + 1] == 0) && (w < 0 || load[c] < load[w])) w = c;
    if (w < 0) return false;
    put(w, segs.np[w][n2u] == 0 ? n2u : n2u + 1, sg, np);
    return true;
  };
  for (; at < two.size(); at++)
    if (!place(two[at], 2)) return -1;
  for (size_t k = 0; k < one.size(); k++)
    if (!place(one[k], 1)) return -1;
  *nt_out = nt;
  *n2u_out = n2u;
  *side_out = side;
  return 0;
}

// The partition as plain arrays [warps][slots] (inspection / CPU tests; no device call).
extern "C" int pcu_gram_wide_plan(int m, int *nt, int *n2u, int *side, int *warps, int *slots,
                                  unsigned char *ti, unsigned char *tj, unsigned char *np) {
  if (!nt || !n2u || !side || !warps || !slots) return 1;
  *warps = PCU_GW_NCW;
  *slots = PCU_GW_MAXSEG;
  if (m < 1 || m > PCU_MAX_COLS) return 1;
  GramSegTable segs;
  bool sd = false;
  if (gram_wide_deal(m, nt, n2u, &sd, &segs)) return 1;
  *side = sd ? 1 : 0;
  if (ti) memcpy(ti, segs.ti, sizeof(segs.ti));
  if (tj) memcpy(tj, segs.tj, sizeof(segs.tj));
  if (np) memcpy(np, segs.np, sizeof(segs.np));
  return 0;
}

static int launch_gram_wide(pcu_ctx *ctx, const ColTable &cols, int m,
                            const double *Dinv, long long n, double *result, int ld,
                            long long *rows_done) {
  int nt = 0, n2u = 0;
  bool side = false;
  GramSegTable segs;
  if (gram_wide_deal(m, &nt, &n2u, &side, &segs)) return -1;
  int stage_bytes = (m + 2) * PCU_GW_COLB;
  stage_bytes = (stage_bytes + 127) / 128 * 128;
  int nstages = (208 * 1024) / stage_bytes;
  if (nstages > PCU_GT_MAXSTAGES) nstages = PCU_GT_MAXSTAGES;
  if (nstages < 2) return -1;
  const long long nslabs = n / PCU_GW_ROWS;
  *rows_done = nslabs * PCU_GW_ROWS;
  const int side_off = side ? (m - 1) * PCU_GW_COLB : -1;
#define PCU_GW_CASE(K)                                                                          \
  case K:                                                                                       \
    return side ? launch_gram_wide_t<K, 1>(ctx, cols, m, nt, segs, Dinv, nslabs, nstages,       \
                                           stage_bytes, result, ld, side_off)                   \
                : launch_gram_wide_t<K, 0>(ctx, cols, m, nt, segs, Dinv, nslabs, nstages,       \
                                           stage_bytes, result, ld, side_off);
  switch (n2u) {
    PCU_GW_CASE(0)
    PCU_GW_CASE(1)
    PCU_GW_CASE(2)
    PCU_GW_CASE(3)
    PCU_GW_CASE(4)
    PCU_GW_CASE(5)
    PCU_GW_CASE(6)
    PCU_GW_CASE(7)
    PCU_GW_CASE(8)
  }
#undef PCU_GW_CASE
  return -1;
}

// General kernel on the row range [lo, hi) (lo a multiple of 64 and of the block
// size), added to `result`.
static int gram_range_accumulate(pcu_ctx *ctx, const ColTable &cols, int m,
                                 const double *Dinv, const double *Cw,
                                 const WDesc &w, long long lo, long long hi,
                                 double *R, int ld, int nt, const double *d2,
                                 int rhs_col) {
  ColTable c2;
  for (int j = 0; j < m; j++) c2.p[j] = cols.p[j] + lo;
  WDesc w2 = w;
  const double *Cw2 = Cw, *d22 = d2;
  if (w.mode == 1) {
    const long long ncon_elems = (long long)w.nwcon * w.nw;
    if (lo >= ncon_elems) {
      w2.mode = 0;
      w2.nwcon = 0;
    } else {
      w2.nwcon = (int)((ncon_elems - lo) / w.nw);
      Cw2 = Cw + lo / w.nw;
      if (d2) d22 = d2 + lo / w.nw;
    }
    w2.wend = (long long)w2.nwcon * w2.wstride;
  }
  const long long n2 = hi - lo;
  switch (nt) {
    case 1: return launch_gram<1, 1, true>(ctx, c2, 0, 0, m, Dinv + lo, Cw2, w2, n2, R, ld, 0, 0, 0, 1, d22, rhs_col);
    case 2: return launch_gram<2, 2, true>(ctx, c2, 0, 0, m, Dinv + lo, Cw2, w2, n2, R, ld, 0, 0, 0, 1, d22, rhs_col);
    case 3: return launch_gram<3, 3, true>(ctx, c2, 0, 0, m, Dinv + lo, Cw2, w2, n2, R, ld, 0, 0, 0, 1, d22, rhs_col);
    case 4: return launch_gram<4, 4, true>(ctx, c2, 0, 0, m, Dinv + lo, Cw2, w2, n2, R, ld, 0, 0, 0, 1, d22, rhs_col);
    default: return launch_gram<5, 5, true>(ctx, c2, 0, 0, m, Dinv + lo, Cw2, w2, n2, R, ld, 0, 0, 0, 1, d22, rhs_col);
  }
}

// Enqueue S = V^T P V into ctx->d_big (col-major, leading dimension *ld_out =
// 8*ceil(m/8); only entries with row >= col are meaningful).
// rhs_col >= 0 (must be m - 1, and m <= 40): that column is the right-hand side
// d1 of a diagonal solve and d2 its constraint part; row rhs_col of S then
// holds V_j . t1 with t1 = D0^-1 (d1, d2)|x instead of V_j . P d1.
int pcu_gram_enqueue(pcu_ctx *ctx, const ColTable &cols, int m,
                     const double *Dinv, const double *Cw, const WDesc &wd,
                     long long n, int *ld_out, const double *d2, int rhs_col) {
  const int nt = (m + 7) / 8;
  const int ld = 8 * nt;
  *ld_out = ld;
  if (m == 0) return 0;
  if (rhs_col >= 0 && rhs_col != m - 1) return 1;
  if (rhs_col >= 0 && nt > 5) {
    // Wide column sets: without weighting constraints the right-hand side is one more
    // plain column (row m - 1 of S is d1^T Dinv V_j = V_j . t1); with them the block
    // part of t1 needs the narrow kernels' correction step.
    if (wd.nwcon > 0) return 1;
    rhs_col = -1;
    d2 = nullptr;
  }
  if (ctx->big_reserve((size_t)ld * ld + 64, 1)) return 1;
  WDesc w = wd;
  if (w.mode == 2 || w.nwcon == 0) {
    w.mode = 0;  // the tensor-core pass computes V^T Dinv V only
  }
  double *R = ctx->d_big;
  int rc = 0;
  const bool fast_ok = nt <= 5 && Dinv != nullptr && n >= 64 &&
                       (w.mode == 0 || (w.mode == 1 && w.nw >= 8));
  const bool tma_ok = nt <= 5 && Dinv != nullptr && n >= 32768 &&
                      (w.mode == 0 || (w.mode == 1 && w.nw == 8)) &&
                      !getenv("PCU_NO_GRAM_TMA");
  if (tma_ok) {
    // bulk-copy staged kernel on the whole slabs, general kernel on the slab that
    // straddles the end of the weighting blocks and on the tail
    long long rows_done = 0, skip_lo = -1;
    int slab_rows = 0;
    rc = launch_gram_tma(ctx, cols, m, Dinv, Cw, w, n, R, ld, d2, rhs_col, nt,
                         w.mode == 0 ? 0 : 8, &rows_done, &skip_lo, &slab_rows);
    if (rc) return 1;
    if (skip_lo >= 0) {
      const long long hi = skip_lo + slab_rows;
      rc = gram_range_accumulate(ctx, cols, m, Dinv, Cw, w, skip_lo, hi, R, ld, nt, d2,
                                 rhs_col);
      if (rc) return rc;
    }
    if (rows_done < n) {
      rc = gram_range_accumulate(ctx, cols, m, Dinv, Cw, w, rows_done, n, R, ld, nt, d2,
                                 rhs_col);
      if (rc) return rc;
    }
  } else if (fast_ok) {
    // straight-line kernel on the clean 64-row chunks, general kernel on the
    // (at most two) ragged ones
    const int nwc = (w.mode == 0) ? 0 : (w.nw == 8 ? 8 : -1);
    rc = launch_gram_fast(ctx, cols, m, Dinv, Cw, w, n, R, ld, nt, nwc, d2, rhs_col);
    if (rc) return rc;
    long long lst[2];
    int nl = 0;
    const long long nfull = n / 64;
    if (w.mode == 1) {
      const long long ncon_elems = (long long)w.nwcon * w.nw;
      const long long cb = ncon_elems / 64;
      if (ncon_elems % 64 != 0 && cb < nfull) lst[nl++] = cb;
    }
    if (n % 64 != 0) lst[nl++] = nfull;
    if (nl > 0) {
      const long long l0 = lst[0], l1 = nl > 1 ? lst[1] : 0;
      switch (nt) {
        case 1: rc = launch_gram<1, 1, true>(ctx, cols, 0, 0, m, Dinv, Cw, w, n, R, ld, l0, l1, nl, 1, d2, rhs_col); break;
        case 2: rc = launch_gram<2, 2, true>(ctx, cols, 0, 0, m, Dinv, Cw, w, n, R, ld, l0, l1, nl, 1, d2, rhs_col); break;
        case 3: rc = launch_gram<3, 3, true>(ctx, cols, 0, 0, m, Dinv, Cw, w, n, R, ld, l0, l1, nl, 1, d2, rhs_col); break;
        case 4: rc = launch_gram<4, 4, true>(ctx, cols, 0, 0, m, Dinv, Cw, w, n, R, ld, l0, l1, nl, 1, d2, rhs_col); break;
        default: rc = launch_gram<5, 5, true>(ctx, cols, 0, 0, m, Dinv, Cw, w, n, R, ld, l0, l1, nl, 1, d2, rhs_col); break;
      }
    }
  } else if (nt <= 5) {
    switch (nt) {
      case 1: rc = launch_gram<1, 1, true>(ctx, cols, 0, 0, m, Dinv, Cw, w, n, R, ld, 0, 0, 0, 0, d2, rhs_col); break;
      case 2: rc = launch_gram<2, 2, true>(ctx, cols, 0, 0, m, Dinv, Cw, w, n, R, ld, 0, 0, 0, 0, d2, rhs_col); break;
      case 3: rc = launch_gram<3, 3, true>(ctx, cols, 0, 0, m, Dinv, Cw, w, n, R, ld, 0, 0, 0, 0, d2, rhs_col); break;
      case 4: rc = launch_gram<4, 4, true>(ctx, cols, 0, 0, m, Dinv, Cw, w, n, R, ld, 0, 0, 0, 0, d2, rhs_col); break;
      default: rc = launch_gram<5, 5, true>(ctx, cols, 0, 0, m, Dinv, Cw, w, n, R, ld, 0, 0, 0, 0, d2, rhs_col); break;
    }
  } else if (nt <= 20 && w.mode == 0 && Dinv != nullptr && n >= 32768 && rhs_col < 0 &&
             !getenv("PCU_NO_GRAM_TMA")) {
    // wide bulk-copy staged kernel (tile pairs dealt to the warps) on the whole
    // 64-row slabs, 40-column block pairs of the general kernel on the tail
    long long rows_done = 0;
    rc = launch_gram_wide(ctx, cols, m, Dinv, n, R, ld, &rows_done);
    if (rc) return 1;
    if (rows_done < n) {
      ColTable c2;
      for (int j = 0; j < m; j++) c2.p[j] = cols.p[j] + rows_done;
      const long long n2 = n - rows_done;
      const int nb = (nt + 4) / 5;
      for (int bi = 0; bi < nb && !rc; bi++) {
        for (int bj = 0; bj <= bi && !rc; bj++) {
          if (bi == bj)
            rc = launch_gram<5, 5, true>(ctx, c2, 40 * bi, 40 * bi, m, Dinv + rows_done, Cw, w, n2, R, ld, 0, 0, 0, 1);
          else
            rc = launch_gram<5, 5, false>(ctx, c2, 40 * bi, 40 * bj, m, Dinv + rows_done, Cw, w, n2, R, ld, 0, 0, 0, 1);
        }
      }
    }
  } else {
    // column blocks of 40; block pairs (bi >= bj); ragged last block is masked
    const int nb = (nt + 4) / 5;
    for (int bi = 0; bi < nb && !rc; bi++) {
      for (int bj = 0; bj <= bi && !rc; bj++) {
        if (bi == bj)
          rc = launch_gram<5, 5, true>(ctx, cols, 40 * bi, 40 * bi, m, Dinv, Cw, w, n, R, ld);
        else
          rc = launch_gram<5, 5, false>(ctx, cols, 40 * bi, 40 * bj, m, Dinv, Cw, w, n, R, ld);
      }
    }
  }
  if (rc) return rc;
  if (wd.mode == 2 && wd.nwcon > 0) {
    const int total = m * m;
    gram_generic_correction<<<(total + 127) / 128, 128, 0, ctx->stream>>>(
        cols, m, Dinv, Cw, wd, R, ld, d2, rhs_col);
    ctx->launches++;
    PCU_CUDA_OK(cudaGetLastError());
  }
  return 0;
}
