// pcu_gram_tma.cuh -- the weighted Gram pass with its operands staged through
// shared memory by the bulk-copy engine (cp.async.bulk + mbarrier ring).
//
// The register-fed kernels (gram_kernel, gram_fast_kernel) can only keep as
// many bytes in flight as their 128-register threads have outstanding loads,
// and stop at ~50 % of HBM bandwidth.  Here one producer warp per CTA streams
// slabs (32 rows per consumer warp) of every column ([A | Z | d1], Dinv, and the Cw / d2 entries of
// the slab's weighting blocks) into a ring of shared-memory stages, one bulk
// copy per column per slab, completion counted on the stage's "full" mbarrier;
// the consumer warps (16 for up to 24 columns, else 8) read their DMMA fragments from shared memory (each warp
// owns 32 rows of the slab = four blocks of 8 rows) and release the stage
// through its "empty" mbarrier.  One CTA per SM, persistent, slabs dealt round
// robin; bytes in flight = (stages - 1) * stage size, independent of registers.
//
// Fragment mapping, block correction and right-hand-side row: exactly as in
// gram_fast_kernel (pcu_gram_fast.cuh).  Only whole slabs that lie entirely
// inside or entirely outside the weighting blocks are handled here; the (at
// most two) remaining row ranges go to gram_kernel.  Included by pcu_gram.cu.
#pragma once

// NCW consumer warps of 32 rows each: a slab has 32 NCW rows
#define PCU_GT_ROWS(NCW) (32 * (NCW))
#define PCU_GT_THREADS(NCW) (32 * ((NCW) + 1))
#define PCU_GT_NPW 4  // producer warps of gram_tma_kernel: one warp instruction serialises its lanes' copies
#define PCU_GT_THREADS_T(NCW) (32 * ((NCW) + PCU_GT_NPW))
#define PCU_GT_COLB(NCW) (PCU_GT_ROWS(NCW) * 8 + 64)  // bytes per staged column (+64: bank shift)
#define PCU_GT_MAXSTAGES 6
#ifndef PCU_GT_RPW3
#define PCU_GT_RPW3 32  // rows per consumer warp of the 17-24 column kernel (32 or 24)
#endif

__device__ __forceinline__ unsigned gt_smem_u32(const void *p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void gt_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void gt_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void gt_mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gt_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "GT_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra GT_DONE_%=;\n"
      "bra GT_WAIT_%=;\n"
      "GT_DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void gt_bulk_g2s(unsigned dst, const void *src,
                                            unsigned bytes, unsigned bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], "
      "%2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// NWC: 0 = no weighting correction, 8 = blocks of exactly 8 rows.
// Slabs [0, slab_con) lie inside the weighting blocks, slab `slab_skip` (if >= 0)
// straddles their end and is skipped, slabs up to nslabs are plain.
// RPW: rows of a slab per consumer warp (32, or 24 = three blocks of 8: smaller
// stages, so that a third one fits for the 22-24 column passes).
template <int NT, int NWC, int NCW, int RPW>
__global__ void __launch_bounds__(PCU_GT_THREADS_T(NCW), 1)
    gram_tma_kernel(const ColTable cols, const int m,
                    const double *__restrict__ Dinv,
                    const double *__restrict__ Cw, const WDesc w,
                    const long long nslabs, const long long slab_con,
                    const long long slab_skip, const int nstages,
                    const int stage_bytes, double *__restrict__ partials,
                    unsigned int *counter, double *__restrict__ result,
                    const int ld, const double *__restrict__ d2,
                    const int rhs_col, const int reverse) {
  // reverse: the slabs are taken from the last one down (the pass before this one,
  // DiagRhsF, walks upwards and leaves its last rows in L2; see NoStreams::REVERSE)
  constexpr int NP = (NT * (NT + 1)) / 2;
  constexpr int ROWS = RPW * NCW, COLB = ROWS * 8 + 64;
  extern __shared__ __align__(128) unsigned char gt_smem[];
  __shared__ __align__(8) unsigned long long gt_full[PCU_GT_MAXSTAGES];
  __shared__ __align__(8) unsigned long long gt_empty[PCU_GT_MAXSTAGES];
  __shared__ double sm[NCW][64];
  __shared__ bool is_last;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gi = lane >> 2, kk = lane & 3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; s++) {
      gt_mbar_init(gt_smem_u32(&gt_full[s]), PCU_GT_NPW);
      gt_mbar_init(gt_smem_u32(&gt_empty[s]), NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const bool with_d2 = (NWC != 0) && (rhs_col >= 0);
  const unsigned col_bytes = ROWS * 8;
  const unsigned blk_bytes = (ROWS / 8) * 8;
  // staged "columns": m vectors, Dinv, then (NWC) Cw and d2 of the slab's blocks
  const int off_dinv = m * COLB;
  const int off_cw = off_dinv + COLB;
  const int off_d2 = off_cw + ROWS;

  double acc[NP][2];
#pragma unroll
  for (int p = 0; p < NP; p++) acc[p][0] = acc[p][1] = 0.0;

  if (warp >= NCW) {
    // ----------------------------------------------------------- producers
    // the m + 3 copies of a slab are dealt over PCU_GT_NPW warps (copy c goes to
    // warp c % NPW), each warp posts the bytes of its share on the full barrier
    const int pw = warp - NCW;
    const int c0 = pw + PCU_GT_NPW * lane;
    unsigned my_plain = 0, my_con = 0;
    for (int c = c0; c < m + 3; c += 32 * PCU_GT_NPW) {
      if (c <= m) {
        my_plain += col_bytes;
        my_con += col_bytes;
      } else if (c == m + 1) {
        my_con += blk_bytes;
      } else if (with_d2) {
        my_con += blk_bytes;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      my_plain += __shfl_xor_sync(0xffffffffu, my_plain, o);
      my_con += __shfl_xor_sync(0xffffffffu, my_con, o);
    }
    long long it = 0;
    for (long long sq = blockIdx.x; sq < nslabs; sq += gridDim.x) {
      const long long slab = reverse ? nslabs - 1 - sq : sq;
      if (slab == slab_skip) continue;
      const int s = (int)(it % nstages);
      const unsigned round = (unsigned)(it / nstages);
      if (round > 0) gt_mbar_wait(gt_smem_u32(&gt_empty[s]), (round - 1) & 1);
      const bool in_con = (NWC != 0) && (slab < slab_con);
      const unsigned full = gt_smem_u32(&gt_full[s]);
      if (lane == 0) gt_mbar_expect_tx(full, in_con ? my_con : my_plain);
      __syncwarp();
      const unsigned base = gt_smem_u32(gt_smem + (size_t)s * stage_bytes);
      const long long row0 = slab * ROWS;
      for (int c = c0; c < m + 3; c += 32 * PCU_GT_NPW) {
        if (c < m) {
          gt_bulk_g2s(base + c * COLB, cols.p[c] + row0, col_bytes, full);
        } else if (c == m) {
          gt_bulk_g2s(base + off_dinv, Dinv + row0, col_bytes, full);
        } else if (in_con && c == m + 1) {
          gt_bulk_g2s(base + off_cw, Cw + row0 / 8, blk_bytes, full);
        } else if (in_con && with_d2 && c == m + 2) {
          gt_bulk_g2s(base + off_d2, d2 + row0 / 8, blk_bytes, full);
        }
      }
      it++;
    }
  } else {
    // ----------------------------------------------------------- consumers
    double cmask = 1.0;
    int colo[NT];  // byte offset of this lane's column of tile t inside a stage
#pragma unroll
    for (int t = 0; t < NT; t++) {
      int c = 8 * t + gi;
      if (c >= m) {  // padded column of the last tile: read a valid one, mask it
        c = m - 1;
        cmask = 0.0;
      }
      colo[t] = c * COLB;
    }
    const bool rhs_lane = (rhs_col >= 0) && (8 * (NT - 1) + gi == rhs_col);
    const double c1 = w.coef_rest;
    const double dc = (kk == 0) ? (w.coef0 - w.coef_rest) : 0.0;
    const int rowb = (warp * RPW + 2 * kk) * 8;  // byte offset of the lane's first row

    long long it = 0;
    for (long long sq = blockIdx.x; sq < nslabs; sq += gridDim.x) {
      const long long slab = reverse ? nslabs - 1 - sq : sq;
      if (slab == slab_skip) continue;
      const int s = (int)(it % nstages);
      const unsigned round = (unsigned)(it / nstages);
      gt_mbar_wait(gt_smem_u32(&gt_full[s]), round & 1);
      const unsigned char *st = gt_smem + (size_t)s * stage_bytes;
      const bool in_con = (NWC != 0) && (slab < slab_con);
      double h[NT], hcw = 0.0, hd = 0.0;
#pragma unroll
      for (int t = 0; t < NT; t++) h[t] = 0.0;
#pragma unroll
      for (int step = 0; step < RPW / 8; step++) {
        const int ro = rowb + step * 64;
        const double2 wv = *reinterpret_cast<const double2 *>(st + off_dinv + ro);
        double2 fb[NT], fa[NT];
#pragma unroll
        for (int t = 0; t < NT; t++) {
          fb[t] = *reinterpret_cast<const double2 *>(st + colo[t] + ro);
          if (t == NT - 1) {
            fb[t].x *= cmask;
            fb[t].y *= cmask;
          }
          fa[t] = make_double2(fb[t].x * wv.x, fb[t].y * wv.y);
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
            // this step is one block of 8 rows: u = sum coef Dinv V over the block
            double u[NT];
#pragma unroll
            for (int t = 0; t < NT; t++) {
              u[t] = fma(dc, fa[t].x, c1 * (fa[t].x + fa[t].y));
              u[t] += shfl_xor_d(u[t], 1);
              u[t] += shfl_xor_d(u[t], 2);
            }
            if (kk == step) {  // park it in k-slot `step`
              const int bi = warp * (RPW / 8) + step;
              hcw = *reinterpret_cast<const double *>(st + off_cw + bi * 8);
              if (rhs_lane) hd = *reinterpret_cast<const double *>(st + off_d2 + bi * 8);
#pragma unroll
              for (int t = 0; t < NT; t++) h[t] = u[t];
            }
          }
        }
      }
      if (NWC != 0) {
        if (in_con) {  // one correction DMMA set for the warp's RPW / 8 blocks
          int pp = 0;
#pragma unroll
          for (int ti = 0; ti < NT; ti++) {
#pragma unroll
            for (int tj = 0; tj <= ti; tj++) {
              dmma884(acc[pp], -hcw * (ti == NT - 1 ? h[ti] - hd : h[ti]), h[tj]);
              pp++;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) gt_mbar_arrive(gt_smem_u32(&gt_empty[s]));
      it++;
    }
  }

  // ---- CTA combine (pair by pair), then grid combine by the last block ----
#pragma unroll
  for (int p = 0; p < NP; p++) {
    if (warp < NCW) {
      sm[warp][gi + 8 * (2 * kk)] = acc[p][0];
      sm[warp][gi + 8 * (2 * kk + 1)] = acc[p][1];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      double v = 0.0;
      for (int ww = 0; ww < NCW; ww++) v += sm[ww][threadIdx.x];
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

// ----------------------------------------------------------------------------
// Wide variant (m up to 160 columns, no weighting correction): FP64-tensor-bound
// (C4: 121 columns, 15 flop/B).  The slab is 64 rows and every consumer warp
// reads ALL of it; the work is split over the lower-triangle TILE PAIRS instead
// of over rows: row ti of the tile triangle is cut into segments of two pairs
// (ti, tj0), (ti, tj0 + 1) that share the weighted A fragment (+ one single-pair
// segment per odd row), and the segments are dealt to the consumer warps so that
// every warp holds N2U two-pair segments plus at most one more two-pair and one
// single-pair segment (nt = 16: 136 pairs = 4 warps x 12 + 8 warps x 11).
//
// 12 consumer warps (three warpgroups, 160 registers each after setmaxnreg) + 4
// producer warps (24 registers) that deal the m + 1 bulk copies of a slab among
// themselves.  ncu of the version with 16 + 1 warps (96 registers, profiles/
// r2t_ncu_raw_C4.csv): a DMMA.8x8x4 held the tensor pipe 20.7 cycles against 16 when
// consecutive DMMAs are independent (the narrow kernel, cuBLAS DGEMM) -- the compiler
// had sunk every fragment load to its use and left the two halves of an accumulator
// two instructions apart.  Here a step works in two chunks of three or four segments:
// the chunk's fragment loads, the scalings, its x-half DMMAs on six to eight different
// accumulators, then the y-half DMMAs.
#define PCU_GW_ROWS 64
#define PCU_GW_NCW 12
#define PCU_GW_NPW 4
#define PCU_GW_COLB (PCU_GW_ROWS * 8 + 64)
#define PCU_GW_MAXN2U 8
#define PCU_GW_MAXSEG (PCU_GW_MAXN2U + 2)  // N2U common segments + 2 optional ones
// setmaxnreg moves registers inside the CTA's own allocation (launch: 128 x 512): the
// producers' 104 x 128 cover the consumers' 32 x 384.  (16 consumer warps + 4 producers
// launch at 96 registers: the consumers could reach 112, not more -- a request beyond the
// pool waits forever.)
#define PCU_GW_CONS_REGS 160
#define PCU_GW_PROD_REGS 24
static_assert((128 - PCU_GW_PROD_REGS) * PCU_GW_NPW >= (PCU_GW_CONS_REGS - 128) * PCU_GW_NCW,
              "setmaxnreg: the producers do not free what the consumers ask for");
static_assert((PCU_GW_NCW + PCU_GW_NPW) * 32 == 512, "gram_wide_kernel is laid out for 16 warps");

struct GramSegTable {
  // segment s of warp w: tile row, first tile column, pairs (1 or 2; 0 = none).
  // Order: the N2U common two-pair segments, then the optional two-pair one, then
  // the optional single-pair one.
  unsigned char ti[PCU_GW_NCW][PCU_GW_MAXSEG];
  unsigned char tj[PCU_GW_NCW][PCU_GW_MAXSEG];
  unsigned char np[PCU_GW_NCW][PCU_GW_MAXSEG];
};

__device__ __forceinline__ double2 gw_lds2(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}

template <int N2U, int SIDE>
__global__ void __launch_bounds__(32 * (PCU_GW_NCW + PCU_GW_NPW), 1)
    gram_wide_kernel(const ColTable cols, const int m, const int nt,
                     const GramSegTable segs, const double *__restrict__ Dinv,
                     const long long nslabs, const int nstages,
                     const int stage_bytes, double *__restrict__ partials,
                     unsigned int *counter, double *__restrict__ result,
                     const int ld, const int reverse, const int side_off) {
  // SIDE: column m - 1 (byte offset side_off in a stage) is alone in its tile -- C4: 120
  // columns + the first solve's right-hand side.  As a 16th tile row it would cost 16 of
  // 136 tile pairs for 121 useful numbers; instead the tiles cover the first m - 1 columns
  // (nt = (m - 1) / 8) and row m - 1 of S is accumulated with DFMAs: warp w takes the
  // columns of tile rows w and w + 12 (straight-line code: a tile row beyond nt reads the
  // zero column), 3 loads + 8 fp64 operations per step next to ~80 DMMAs.
  constexpr int NSEG = N2U + 2;
  constexpr int CH = NSEG <= 8 ? (NSEG + 1) / 2 : 4;  // segments per chunk of a step
  constexpr int NCT = PCU_GW_NCW * 32;  // consumer threads
  extern __shared__ __align__(128) unsigned char gt_smem[];
  __shared__ __align__(8) unsigned long long gt_full[PCU_GT_MAXSTAGES];
  __shared__ __align__(8) unsigned long long gt_empty[PCU_GT_MAXSTAGES];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gi = lane >> 2, kk = lane & 3;
  const int npairs_tot = nt * (nt + 1) / 2;
  const size_t pstride = (size_t)npairs_tot * 64 + (SIDE ? (size_t)ld : 0);  // per CTA

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; s++) {
      gt_mbar_init(gt_smem_u32(&gt_full[s]), PCU_GW_NPW);
      gt_mbar_init(gt_smem_u32(&gt_empty[s]), PCU_GW_NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned col_bytes = PCU_GW_ROWS * 8;
  const int off_dinv = m * PCU_GW_COLB;
  const int off_zero = off_dinv + PCU_GW_COLB;  // all-zero column: padding of the last tile
  for (int i = threadIdx.x; i < nstages * (PCU_GW_ROWS * 8 / 16); i += blockDim.x) {
    const int s = i / (PCU_GW_ROWS * 8 / 16), o = i % (PCU_GW_ROWS * 8 / 16);
    *reinterpret_cast<double2 *>(gt_smem + (size_t)s * stage_bytes + off_zero + 16 * o) =
        make_double2(0.0, 0.0);
  }
  __syncthreads();

  if (warp >= PCU_GW_NCW) {
    // ----------------------------------------------------------- producers
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PCU_GW_PROD_REGS));
    const int pw = warp - PCU_GW_NCW;
    const int c0 = pw + PCU_GW_NPW * lane;  // copy c goes to warp c % NPW
    unsigned mine = 0;
    for (int c = c0; c <= m; c += 32 * PCU_GW_NPW) mine += col_bytes;
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    long long it = 0;
    for (long long sq = blockIdx.x; sq < nslabs; sq += gridDim.x, it++) {
      const long long slab = reverse ? nslabs - 1 - sq : sq;
      const int s = (int)(it % nstages);
      const unsigned round = (unsigned)(it / nstages);
      if (round > 0) gt_mbar_wait(gt_smem_u32(&gt_empty[s]), (round - 1) & 1);
      const unsigned full = gt_smem_u32(&gt_full[s]);
      if (lane == 0) gt_mbar_expect_tx(full, mine);
      __syncwarp();
      const unsigned base = gt_smem_u32(gt_smem + (size_t)s * stage_bytes);
      const long long row0 = slab * PCU_GW_ROWS;
      for (int c = c0; c <= m; c += 32 * PCU_GW_NPW) {
        if (c < m) gt_bulk_g2s(base + c * PCU_GW_COLB, cols.p[c] + row0, col_bytes, full);
        else gt_bulk_g2s(base + off_dinv, Dinv + row0, col_bytes, full);
      }
    }
    return;
  }
  // ------------------------------------------------------------- consumers
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PCU_GW_CONS_REGS));
  double acc[NSEG][2][2];
#pragma unroll
  for (int s = 0; s < NSEG; s++)
    acc[s][0][0] = acc[s][0][1] = acc[s][1][0] = acc[s][1][1] = 0.0;
  {
    // this lane's column offsets per segment (padded columns read the zero column)
    unsigned offA[NSEG], offB0[NSEG], offB1[NSEG];
    auto colof = [&](int t) -> unsigned {
      const int c = 8 * t + gi;
      return (unsigned)(c < m ? c * PCU_GW_COLB : off_zero);
    };
#pragma unroll
    for (int s = 0; s < NSEG; s++) {
      offA[s] = colof(segs.ti[warp][s]);
      offB0[s] = colof(segs.tj[warp][s]);
      offB1[s] = colof(segs.tj[warp][s] + 1);
    }
    // pairs of slot s: 2 for the common segments (compile time), 0..2 / 0..1 for the
    // two optional ones (warp-uniform)
    const int npx = segs.np[warp][N2U], npy = segs.np[warp][N2U + 1];
    double sacc[2] = {0.0, 0.0}, srr = 0.0;
    const unsigned offS0 = colof(warp), offS1 = colof(warp + PCU_GW_NCW);
    const bool own0 = warp < nt, own1 = warp + PCU_GW_NCW < nt;
    long long it = 0;
    for (long long sq = blockIdx.x; sq < nslabs; sq += gridDim.x, it++) {
      const int s = (int)(it % nstages);
      const unsigned round = (unsigned)(it / nstages);
      gt_mbar_wait(gt_smem_u32(&gt_full[s]), round & 1);
      const unsigned sb = gt_smem_u32(gt_smem) + (unsigned)s * (unsigned)stage_bytes;
#pragma unroll 2
      for (int step = 0; step < PCU_GW_ROWS / 8; step++) {
        const unsigned ro = sb + (unsigned)(step * 64 + kk * 16);
        const double2 wv = gw_lds2(ro + (unsigned)off_dinv);
        if constexpr (SIDE) {
          const double2 rr = gw_lds2(ro + (unsigned)side_off);
          const double2 c0v = gw_lds2(ro + (own0 ? offS0 : (unsigned)off_zero));
          const double2 c1v = gw_lds2(ro + (own1 ? offS1 : (unsigned)off_zero));
          const double rwx = rr.x * wv.x, rwy = rr.y * wv.y;
          sacc[0] = fma(c0v.x, rwx, sacc[0]);
          sacc[0] = fma(c0v.y, rwy, sacc[0]);
          sacc[1] = fma(c1v.x, rwx, sacc[1]);
          sacc[1] = fma(c1v.y, rwy, sacc[1]);
          srr = fma(rr.x, rwx, srr);
          srr = fma(rr.y, rwy, srr);
        }
#pragma unroll
        for (int c0 = 0; c0 < NSEG; c0 += CH) {
          double2 a[CH], b0[CH], b1[CH];
#pragma unroll
          for (int u = 0; u < CH; u++) {
            const int sg = c0 + u;
            a[u] = b0[u] = b1[u] = make_double2(0.0, 0.0);
            if (sg < N2U) {
              a[u] = gw_lds2(ro + offA[sg]);
              b0[u] = gw_lds2(ro + offB0[sg]);
              b1[u] = gw_lds2(ro + offB1[sg]);
            } else if (sg < NSEG) {
              const int np = sg == N2U ? npx : npy;
              if (np > 0) {
                a[u] = gw_lds2(ro + offA[sg]);
                b0[u] = gw_lds2(ro + offB0[sg]);
              }
              if (np > 1) b1[u] = gw_lds2(ro + offB1[sg]);
            }
          }
#pragma unroll
          for (int u = 0; u < CH; u++) {
            if (c0 + u < NSEG) {
              a[u].x *= wv.x;
              a[u].y *= wv.y;
            }
          }
#pragma unroll
          for (int h = 0; h < 2; h++) {  // x halves of the chunk, then its y halves
#pragma unroll
            for (int u = 0; u < CH; u++) {
              const int sg = c0 + u;
              const double av = h == 0 ? a[u].x : a[u].y;
              const double bv0 = h == 0 ? b0[u].x : b0[u].y;
              const double bv1 = h == 0 ? b1[u].x : b1[u].y;
              if (sg < N2U) {
                dmma884(acc[sg][0], av, bv0);
                dmma884(acc[sg][1], av, bv1);
              } else if (sg < NSEG) {
                const int np = sg == N2U ? npx : npy;
                if (np > 0) dmma884(acc[sg][0], av, bv0);
                if (np > 1) dmma884(acc[sg][1], av, bv1);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) gt_mbar_arrive(gt_smem_u32(&gt_empty[s]));
    }
    // every tile pair is owned by exactly one warp of the CTA
#pragma unroll
    for (int sg = 0; sg < NSEG; sg++) {
      const int npv = segs.np[warp][sg];
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (q < npv) {
          const int ti = segs.ti[warp][sg], tj = segs.tj[warp][sg] + q;
          const int p = ti * (ti + 1) / 2 + tj;
          double *dst = partials + (size_t)blockIdx.x * pstride + (size_t)p * 64;
          dst[gi + 8 * (2 * kk)] = acc[sg][q][0];
          dst[gi + 8 * (2 * kk + 1)] = acc[sg][q][1];
        }
      }
    }
    if constexpr (SIDE) {
      // the four lanes of a column hold two of a step's eight rows each
#pragma unroll
      for (int q = 0; q < 2; q++) {
        double v = sacc[q];
        v += shfl_xor_d(v, 1);
        v += shfl_xor_d(v, 2);
        const int ti = warp + PCU_GW_NCW * q;
        if (kk == 0 && ti < nt)
          partials[(size_t)blockIdx.x * pstride + (size_t)npairs_tot * 64 + 8 * ti + gi] = v;
      }
      double v = srr;  // the side column with itself (warp 0, lanes 0..3)
      v += shfl_xor_d(v, 1);
      v += shfl_xor_d(v, 2);
      if (warp == 0 && lane == 0)
        partials[(size_t)blockIdx.x * pstride + (size_t)npairs_tot * 64 + (m - 1)] = v;
    }
  }
  // the consumer threads alone from here on (the producers have returned)
  __threadfence();
  asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
  if (is_last) {
    __threadfence();
    for (int idx = threadIdx.x; idx < npairs_tot * 64; idx += NCT) {
      const double v = pcu_ordered_sum(partials + idx, pstride, 0u, 1u, gridDim.x);
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
    if constexpr (SIDE) {  // row m - 1 of S
      for (int c = threadIdx.x; c < m; c += NCT)
        result[(size_t)(m - 1) + (size_t)ld * c] = pcu_ordered_sum(
            partials + (size_t)npairs_tot * 64 + c, pstride, 0u, 1u, gridDim.x);
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}
