// pcu_gram_fast.cuh -- straight-line DMMA Gram kernel for the common case
// (one diagonal column block, m <= 40, no weighting constraints or aligned
// blocks of nw >= 8 rows).  Same fragment mapping as gram_kernel (pcu_gram.cu)
// but the 64-row chunk is fully unrolled: per-step addressing is immediate
// offsets from four running pointers, loads are double-buffered in groups of
// two 8-row steps (no register moves), and chunks that are ragged (tail of the
// vector, or straddling the end of the weighting constraints) are left to the
// general kernel.  Included by pcu_gram.cu only.
#pragma once

#ifndef PCU_GRAM_G
#define PCU_GRAM_G 2   // 8-row steps per load group (double-buffered)
#endif
#ifndef PCU_GRAM_MINB
#define PCU_GRAM_MINB 2
#endif

template <int NT>
struct GramBuf {
  double2 wv[PCU_GRAM_G];
  double2 f[PCU_GRAM_G][NT];
};

// NWC: 0 = no weighting correction, 8 = blocks of exactly 8 rows,
//     -1 = runtime block size w.nw in {16, 32, 64}
template <int NT, int NWC>
__global__ void __launch_bounds__(PCU_THREADS, PCU_GRAM_MINB)
    gram_fast_kernel(const ColTable cols, const int m,
                     const double *__restrict__ Dinv,
                     const double *__restrict__ Cw, const WDesc w,
                     const long long n, double *__restrict__ partials,
                     unsigned int *counter, double *__restrict__ result,
                     const int ld, const double *__restrict__ d2,
                     const int rhs_col) {
  // rhs_col >= 0: column rhs_col (the last one) is the right-hand side d1 of a
  // diagonal solve with constraint part d2; its ROW of the result then holds
  // V_j . t1, t1 = D0^-1 (d1, d2)|x  (SM.cpp:160-190): the block sums u of that
  // column are shifted by -d2 on the A side of the correction.
  constexpr int NP = (NT * (NT + 1)) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gi = lane >> 2, kk = lane & 3;
  const int nwarps_cta = blockDim.x >> 5;
  const long long gwarp = (long long)blockIdx.x * nwarps_cta + warp;
  const long long nwarps = (long long)gridDim.x * nwarps_cta;

  double acc[NP][2];
#pragma unroll
  for (int p = 0; p < NP; p++) acc[p][0] = acc[p][1] = 0.0;
  double u[NT], h[NT], hcw = 0.0, hd = 0.0;
#pragma unroll
  for (int t = 0; t < NT; t++) u[t] = h[t] = 0.0;
  const bool rhs_lane = (rhs_col >= 0) && (8 * (NT - 1) + gi == rhs_col);

  const double *col[NT];
  double cmask = 1.0;
#pragma unroll
  for (int t = 0; t < NT; t++) {
    int c = 8 * t + gi;
    if (c >= m) {  // padded column of the last tile: reload a valid column, mask it
      c = m - 1;
      cmask = 0.0;
    }
    col[t] = cols.p[c] + 2 * kk;
  }
  const double *dinv = Dinv + 2 * kk;

  const int nw = (NWC == 8) ? 8 : w.nw;
  const long long ncon_elems = (NWC != 0) ? (long long)w.nwcon * nw : 0;
  // coefficient of row r: coef0 at block starts, coef_rest elsewhere.  Only lane
  // kk == 0 can hold a block start (its .x row), and only on steps that begin a
  // block.
  const double c1 = w.coef_rest;
  const double dc = (kk == 0) ? (w.coef0 - w.coef_rest) : 0.0;

  const long long nfull = n / 64;  // chunks whose 64 rows all exist
  auto clean = [&](long long chunk) -> bool {
    if (NWC == 0) return true;
    const long long lo = chunk * 64;
    return (lo + 64 <= ncon_elems) || (lo >= ncon_elems);
  };

  auto load_group = [&](long long chunk, int g, GramBuf<NT> &b) {
    const long long off = chunk * 64 + g * (8 * PCU_GRAM_G);
#pragma unroll
    for (int s = 0; s < PCU_GRAM_G; s++) {
      b.wv[s] = *reinterpret_cast<const double2 *>(dinv + off + 8 * s);
#pragma unroll
      for (int t = 0; t < NT; t++)
        b.f[s][t] = *reinterpret_cast<const double2 *>(col[t] + off + 8 * s);
    }
  };

  auto compute_group = [&](long long chunk, int g, const GramBuf<NT> &b,
                           bool in_con) {
#pragma unroll
    for (int s = 0; s < PCU_GRAM_G; s++) {
      const int step = PCU_GRAM_G * g + s;
      double2 fb[NT], fa[NT];
#pragma unroll
      for (int t = 0; t < NT; t++) {
        fb[t] = b.f[s][t];
        if (t == NT - 1) {
          fb[t].x *= cmask;
          fb[t].y *= cmask;
        }
        fa[t] = make_double2(fb[t].x * b.wv[s].x, fb[t].y * b.wv[s].y);
      }
      int p = 0;
#pragma unroll
      for (int ti = 0; ti < NT; ti++) {
#pragma unroll
        for (int tj = 0; tj <= ti; tj++) {
          dmma884(acc[p], fa[ti].x, fb[tj].x);
          p++;
        }
      }
      p = 0;
#pragma unroll
      for (int ti = 0; ti < NT; ti++) {
#pragma unroll
        for (int tj = 0; tj <= ti; tj++) {
          dmma884(acc[p], fa[ti].y, fb[tj].y);
          p++;
        }
      }
      if (NWC != 0) {
        if (in_con) {
          const bool bstart = (NWC == 8) ? true : (((step * 8) & (nw - 1)) == 0);
          const bool bend = (NWC == 8) ? true : ((((step + 1) * 8) & (nw - 1)) == 0);
#pragma unroll
          for (int t = 0; t < NT; t++) {
            double v = c1 * (fa[t].x + fa[t].y);
            if (bstart) v = fma(dc, fa[t].x, v);
            u[t] += v;
          }
          if (bend) {
#pragma unroll
            for (int t = 0; t < NT; t++) {
              u[t] += shfl_xor_d(u[t], 1);
              u[t] += shfl_xor_d(u[t], 2);
            }
            const long long ci = (chunk * 64 + step * 8) / nw;
            const int blk = (int)(ci & 3);
            if (kk == blk) {
              hcw = Cw[ci];
#pragma unroll
              for (int t = 0; t < NT; t++) h[t] = u[t];
              hd = rhs_lane ? d2[ci] : 0.0;
            }
#pragma unroll
            for (int t = 0; t < NT; t++) u[t] = 0.0;
            if (blk == 3 || step == 7) {
              int pp = 0;
#pragma unroll
              for (int ti = 0; ti < NT; ti++) {
#pragma unroll
                for (int tj = 0; tj <= ti; tj++) {
                  dmma884(acc[pp], -hcw * (ti == NT - 1 ? h[ti] - hd : h[ti]), h[tj]);
                  pp++;
                }
              }
              hcw = 0.0;
              hd = 0.0;
#pragma unroll
              for (int t = 0; t < NT; t++) h[t] = 0.0;
            }
          }
        }
      }
    }
  };

  // first clean chunk of this warp
  long long chunk = gwarp;
  while (chunk < nfull && !clean(chunk)) chunk += nwarps;
  GramBuf<NT> ba, bb;
  if (chunk < nfull) load_group(chunk, 0, ba);
  while (chunk < nfull) {
    const bool in_con = (NWC != 0) && (chunk * 64 < ncon_elems);
    long long next = chunk + nwarps;
    while (next < nfull && !clean(next)) next += nwarps;
    constexpr int NG = 8 / PCU_GRAM_G;  // groups per chunk (even)
#pragma unroll
    for (int g = 0; g < NG; g += 2) {
      load_group(chunk, g + 1, bb);
      compute_group(chunk, g, ba, in_con);
      if (g + 2 < NG) load_group(chunk, g + 2, ba);
      else if (next < nfull) load_group(next, 0, ba);
      compute_group(chunk, g + 1, bb, in_con);
    }
    chunk = next;
  }

  // ---- CTA combine (pair by pair), then grid combine by the last block ----
  __shared__ double sm[PCU_THREADS / 32][64];
  __shared__ bool is_last;
#pragma unroll
  for (int p = 0; p < NP; p++) {
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
      const int p = idx >> 6, e = idx & 63;
      int ti = 0, q = p;
      while (q > ti) {
        q -= ti + 1;
        ti++;
      }
      const int tj = q;
      const int row = 8 * ti + (e & 7);
      const int cc = 8 * tj + (e >> 3);
      if (row < ld && cc < ld) result[(size_t)row + (size_t)ld * cc] = v;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}
