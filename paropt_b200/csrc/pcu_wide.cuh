// pcu_wide.cuh -- bulk-copy staged harness for the pass-2 kernels of the KKT solve
// with MANY columns (33..160 constraint gradients + quasi-Newton vectors; C4: 120)
// and no weighting constraints (any even number of rows).
//
// tma_tile_kernel gives every tile to a group of two warps and every element pair to
// one lane, which then walks ALL columns: with 130+ streams only 64-row tiles fit the
// shared memory three times, three tiles keep three warps busy, and the running dot
// products [A|Z]^T t1' of the fused refinement half-solve (Pass2R1F) would need 120+
// accumulators per thread.  Here a tile belongs to a TEAM of six consumer warps that
// split the COLUMNS of the shared-memory tile instead of its rows:
//   phase A  warp w of the team: columns w, w + 6, ... x all 64 rows (two per lane,
//            128-bit conflict-free loads) -> partial sums of  d1 + V alpha,
//            V beta (the quasi-Newton / constraint part of the residual row) and
//            A alpha (the `A p_z` hand-over) into the team's scratch
//   phase E  the team's first warp adds the six partials in warp order, patches the
//            d1 slot of the stage in place, fills the `lin` slot, stores A p_z, and
//            runs the functor's ordinary phases (ncols = 0) on the 64 rows
//   phase F  (DOTS) every warp: its columns x the t1' the functor left in the stage
//            -> running dot products in registers (<= 27 per lane)
// Two teams work on alternate tiles, so the single-warp phase E of one overlaps the
// column phases of the other; the ring holds three 64-row stages (66-68 KB each at
// C4) filled by four producer warps with one cp.async.bulk per stream.  Reference
// functions: the same as Pass2R1F / Pass2SF (IP.cpp:2074-2369, 1337-1583, 2700-2737).
// All sums have a fixed order (columns within a warp, warps within a team, teams,
// CTAs): results are bit-reproducible.
#pragma once

#include "pcu_common.cuh"
#include "pcu_ctx.cuh"

#define PCU_WT_ROWS 64
#define PCU_WT_TEAMS 2
#define PCU_WT_TEAMW 6   // warps per team

struct WidePlan {
  long long ntiles;  // tiles, the last one possibly partial
  int tail;          // rows of the last tile when it is partial (even), else 0
  int nstages, stage_bytes, npw, col_base;
  unsigned long long nmap[3];
  unsigned noff[24];
  unsigned scr_off;  // the teams' scratch, after the ring
  int m, nca;        // columns; the first nca of them also go into apz
  int reverse;       // F::REVERSE: tiles from the last one down (see NoStreams::REVERSE)
};

// producer-side visitor: the functor's fixed N-streams (W-streams do not exist here)
template <int NFIX>
struct WideAssign {
  unsigned long long nmap[3];
  int col_base;
  int me_warp, me_lane, c, npw;
  int cnt;
  const double *p0, *p1;
  unsigned off0, off1;
  __device__ __forceinline__ void take(const double *ptr, unsigned off) {
    if (ptr == nullptr) return;
    const int cc = c++;
    if (cc % npw != me_warp || (cc / npw) % 32 != me_lane) return;
    if (cnt == 0) {
      p0 = ptr; off0 = off;
    } else {
      p1 = ptr; off1 = off;
    }
    cnt++;
  }
  __device__ __forceinline__ void n(int slot, const double *ptr) {
    const int cs = slot < NFIX ? (int)((nmap[(slot >> 3) % 3] >> ((slot & 7) * 8)) & 0xffull)
                               : col_base + (slot - NFIX);
    take(ptr, (unsigned)cs * (PCU_WT_ROWS * 8));
  }
  __device__ __forceinline__ void w(int, const double *) {}
};

__device__ __forceinline__ double2 wt_lds2(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void wt_sts2(unsigned a, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void wt_team_sync(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(2 + team), "n"(PCU_WT_TEAMW * 32) : "memory");
}

// F: Pass2R1F<0, 1> (DOTS = 1: lin slot, t slot, dot products) or Pass2SF (DOTS = 0).
// f.ncols == 0 and f.apz == nullptr; the columns are f.V.p[0 .. plan.m), their
// coefficients f.alpha (and f.beta with DOTS).
template <class F, int DOTS, int MAXJ>
__global__ void __launch_bounds__(PCU_TMA_MAXWARPS * 32, 1)
    wide_tile_kernel(const F f, const RedBuf rb, const WidePlan plan, double *apz,
                     double *dot_partials, unsigned int *dot_counter, double *dot_result) {
  constexpr int ROWS = PCU_WT_ROWS, TEAMW = PCU_WT_TEAMW;
  constexpr int NV = DOTS ? 3 : 2;  // partial sums per row: d, (lin,) A p_z
  constexpr int NCW = PCU_WT_TEAMS * TEAMW;
  extern __shared__ double2 pcu_dyn_smem[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(pcu_dyn_smem);
  __shared__ __align__(8) unsigned long long wt_full[PCU_TMA_MAXSTAGES];
  __shared__ __align__(8) unsigned long long wt_empty[PCU_TMA_MAXSTAGES];
  __shared__ double wt_dots[DOTS ? PCU_WT_TEAMS : 1][DOTS ? PCU_MAX_COLS : 1];
  __shared__ bool wt_last;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = plan.nstages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) {
      tt_mbar_init(tt_smem_u32(&wt_full[s]), plan.npw);
      tt_mbar_init(tt_smem_u32(&wt_empty[s]), TEAMW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned full0 = tt_smem_u32(wt_full), empty0 = tt_smem_u32(wt_empty);
  const unsigned smem0 = tt_smem_u32(smem);
  const int ntl =
      (long long)blockIdx.x < plan.ntiles ? (int)((plan.ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  const int m = plan.m;

  if (warp < PCU_TMA_NPW) {
    // ------------------------------------------------- producers (warpgroup 0)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PCU_TMA_PROD_REGS));
    if (warp >= plan.npw) return;
    WideAssign<F::NFIX> as;
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
    f.tstreams(as);
    for (int j = 0; j < m; j++) as.n(F::NFIX + j, f.V.p[j]);
    unsigned ncopies = (unsigned)as.cnt;
    for (int o = 16; o > 0; o >>= 1) ncopies += __shfl_xor_sync(0xffffffffu, ncopies, o);
    int s = 0;
    unsigned round = 0;
    for (int kt = 0; kt < ntl; kt++) {
      const long long tq = blockIdx.x + (long long)kt * gridDim.x;
      const long long tile = plan.reverse ? plan.ntiles - 1 - tq : tq;
      // the ragged last tile: only its rows are copied (an even number: 16-byte units)
      const unsigned nbytes =
          (plan.tail > 0 && tile == plan.ntiles - 1) ? (unsigned)plan.tail * 8u : ROWS * 8u;
      if (round > 0) tt_mbar_wait(empty0 + 8u * s, (round - 1) & 1);
      const unsigned full = full0 + 8u * s;
      if (lane == 0) tt_mbar_expect_tx(full, ncopies * nbytes);
      __syncwarp();
      const unsigned base = smem0 + (unsigned)s * (unsigned)plan.stage_bytes;
      if (as.cnt > 0) tt_bulk_g2s(base + as.off0, as.p0 + tile * ROWS, nbytes, full);
      if (as.cnt > 1) tt_bulk_g2s(base + as.off1, as.p1 + tile * ROWS, nbytes, full);
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
  const int cw = warp - PCU_TMA_NPW;
  const int team = cw / TEAMW, tw = cw % TEAMW;
  const unsigned scr = smem0 + plan.scr_off + (unsigned)team * (TEAMW * NV * ROWS * 8u);
  const unsigned col0 = (unsigned)plan.col_base * (ROWS * 8u);
  double dacc[DOTS ? MAXJ : 1];
#pragma unroll
  for (int jj = 0; jj < (DOTS ? MAXJ : 1); jj++) dacc[jj] = 0.0;
  WDesc w0;
  w0.nwcon = 0;
  w0.mode = 0;
  w0.nw = 1;
  w0.nw_log2 = 0;
  w0.wstart = w0.wend = 0;
  w0.wstride = 1;
  w0.coef0 = w0.coef_rest = w0.wconst = 0.0;

  for (int kt = team; kt < ntl; kt += PCU_WT_TEAMS) {
    const long long tq = blockIdx.x + (long long)kt * gridDim.x;
    const long long tile = plan.reverse ? plan.ntiles - 1 - tq : tq;
    const int s = kt % S;
    // rows of this tile; lanes beyond them neither store nor reduce (their part of the
    // stage holds leftovers of an earlier tile: finite numbers nobody looks at)
    const int rows = (plan.tail > 0 && tile == plan.ntiles - 1) ? plan.tail : ROWS;
    const bool live = 2 * lane < rows;
    tt_mbar_wait(full0 + 8u * s, (unsigned)(kt / S) & 1u);
    const unsigned base = smem0 + (unsigned)s * (unsigned)plan.stage_bytes;
    // ---- phase A: this warp's columns x 64 rows
    {
      double2 pd = make_double2(0.0, 0.0), pl = pd, pq = pd;
      const unsigned cb = base + col0 + (unsigned)lane * 16u;
      for (int j = tw; j < m; j += 2 * TEAMW) {
        const int j2 = j + TEAMW;
        const bool has2 = j2 < m;
        const double2 c0 = wt_lds2(cb + (unsigned)j * (ROWS * 8u));
        double2 c1 = make_double2(0.0, 0.0);
        if (has2) c1 = wt_lds2(cb + (unsigned)j2 * (ROWS * 8u));
        const double a0 = f.alpha.v[j], a1 = has2 ? f.alpha.v[j2] : 0.0;
        pd.x = fma(a0, c0.x, pd.x);
        pd.y = fma(a0, c0.y, pd.y);
        if (j < plan.nca) {
          pq.x = fma(a0, c0.x, pq.x);
          pq.y = fma(a0, c0.y, pq.y);
        }
        pd.x = fma(a1, c1.x, pd.x);
        pd.y = fma(a1, c1.y, pd.y);
        if (j2 < plan.nca) {
          pq.x = fma(a1, c1.x, pq.x);
          pq.y = fma(a1, c1.y, pq.y);
        }
        if constexpr (DOTS) {
          const double b0 = f.beta.v[j], b1 = has2 ? f.beta.v[j2] : 0.0;
          pl.x = fma(b0, c0.x, pl.x);
          pl.y = fma(b0, c0.y, pl.y);
          pl.x = fma(b1, c1.x, pl.x);
          pl.y = fma(b1, c1.y, pl.y);
        }
      }
      const unsigned p = scr + (unsigned)(tw * NV) * (ROWS * 8u) + (unsigned)lane * 16u;
      wt_sts2(p, pd);
      wt_sts2(p + ROWS * 8u, pq);
      if constexpr (DOTS) wt_sts2(p + 2u * ROWS * 8u, pl);
    }
    wt_team_sync(team);
    // ---- phase E: the team's first warp finishes the rows
    double2 sd = make_double2(0.0, 0.0), sq = sd, sl = sd;
    if (tw == 0) {
#pragma unroll
      for (int w2 = 0; w2 < TEAMW; w2++) {
        const unsigned p = scr + (unsigned)(w2 * NV) * (ROWS * 8u) + (unsigned)lane * 16u;
        const double2 a = wt_lds2(p), b = wt_lds2(p + ROWS * 8u);
        sd.x += a.x;
        sd.y += a.y;
        sq.x += b.x;
        sq.y += b.y;
        if constexpr (DOTS) {
          const double2 c = wt_lds2(p + 2u * ROWS * 8u);
          sl.x += c.x;
          sl.y += c.y;
        }
      }
    }
    // without phase F the other warps go on to the team's next tile: the scratch may be
    // overwritten as soon as the first warp holds the sums
    if constexpr (!DOTS) wt_team_sync(team);
    if (tw == 0) {
      const unsigned ad = base + plan.noff[F::S_D1] + (unsigned)lane * 16u;
      double2 dv = wt_lds2(ad);
      dv.x += sd.x;
      dv.y += sd.y;
      wt_sts2(ad, dv);
      if constexpr (DOTS) wt_sts2(base + plan.noff[F::S_LIN] + (unsigned)lane * 16u, sl);
      if (apz && live) *reinterpret_cast<double2 *>(apz + tile * ROWS + 2 * lane) = sq;
      __syncwarp();
      SSrc<ROWS, F::NFIX> src;
      src.nb = base;
      src.wb = base;
      src.row0 = tile * ROWS;
      src.con0 = 0;
      src.wpitch = 0;
#pragma unroll
      for (int q = 0; q < F::NFIX; q++) src.off[q] = plan.noff[q];
      src.col0 = col0;
      if (live) tile_pair(f, src, w0, src.row0 + 2 * lane, 0ll, 1, acc);
    }
    if constexpr (DOTS) {
      wt_team_sync(team);
      // ---- phase F: this warp's columns against the t1' of the rows
      if (live) {  // (the rows beyond a ragged tile may hold anything, NaNs included)
        const double2 t = wt_lds2(base + plan.noff[F::S_T] + (unsigned)lane * 16u);
        const unsigned cb = base + col0 + (unsigned)lane * 16u;
#pragma unroll
        for (int jj = 0; jj < MAXJ; jj++) {
          const int j = tw + TEAMW * jj;
          if (j < m) {
            const double2 c = wt_lds2(cb + (unsigned)j * (ROWS * 8u));
            dacc[jj] = fma(t.x, c.x, dacc[jj]);
            dacc[jj] = fma(t.y, c.y, dacc[jj]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) tt_mbar_arrive(empty0 + 8u * s);
  }

  const int ctid = (int)threadIdx.x - PCU_TMA_NPW * 32;  // consumer thread index
  if constexpr (DOTS) {
    // warp totals -> team totals -> CTA partials -> the last CTA adds them in order
#pragma unroll
    for (int jj = 0; jj < MAXJ; jj++) {
      const int j = tw + TEAMW * jj;
      double v = dacc[jj];
      for (int o = 16; o > 0; o >>= 1) v += shfl_down_d(v, o);
      if (lane == 0 && j < m) wt_dots[team][j] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
    for (int j = ctid; j < m; j += NCW * 32)
      dot_partials[(size_t)blockIdx.x * m + j] = wt_dots[0][j] + wt_dots[1][j];
    __threadfence();
    asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
    if (ctid == 0) {
      const unsigned int t = atomicAdd(dot_counter, 1u);
      wt_last = (t == gridDim.x - 1);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
    if (wt_last) {
      __threadfence();
      for (int j = cw; j < m; j += NCW) {
        double v = pcu_ordered_sum(dot_partials + j, (size_t)m, (unsigned)lane, 32u, gridDim.x);
        for (int o = 16; o > 0; o >>= 1) v += shfl_down_d(v, o);
        if (lane == 0) dot_result[j] = v;
      }
      if (ctid == 0) *dot_counter = 0u;
    }
  }
  if (F::NS + F::NX + F::NM > 0) {
    f.finalize(acc);
    finish_reduction<F::NS, F::NX, F::NM, AT, NCW>(acc, rb, ctid, NCW * 32);
  }
}

// Host side.  Returns -1 when the launch does not qualify (the caller then takes the
// register-fed kernels), 1 on error.  On success with DOTS the m dot products are in
// ctx->d_big (ctx->big_fetch(m, ...)).
template <class F, int DOTS, int MAXJ>
int pcu_launch_wide_t(pcu_ctx *ctx, F f, long long n, const WDesc &w, RedBuf rb, int m);

// MAXJ: dot products per lane -- 22 up to 132 columns (C4: 120), 27 up to PCU_MAX_COLS
template <class F, int DOTS>
int pcu_launch_wide(pcu_ctx *ctx, F f, long long n, const WDesc &w, RedBuf rb, int m) {
  if (DOTS && m > PCU_WT_TEAMW * 22) return pcu_launch_wide_t<F, DOTS, DOTS ? 27 : 1>(ctx, f, n, w, rb, m);
  return pcu_launch_wide_t<F, DOTS, DOTS ? 22 : 1>(ctx, f, n, w, rb, m);
}

template <class F, int DOTS, int MAXJ>
int pcu_launch_wide_t(pcu_ctx *ctx, F f, long long n, const WDesc &w, RedBuf rb, int m) {
  constexpr int ROWS = PCU_WT_ROWS;
  static_assert(F::NFIX <= 24, "fixed slots");
  if (ctx->no_tma_tile || w.nwcon > 0 || m < 1 || m > PCU_MAX_COLS) return -1;
  if (DOTS && m > PCU_WT_TEAMW * MAXJ) return -1;
  if (n & 1) return -1;  // the ragged last tile is copied in 16-byte units
  WidePlan plan;
  plan.tail = (int)(n % ROWS);
  plan.ntiles = n / ROWS + (plan.tail > 0 ? 1 : 0);
  if (plan.ntiles < (ctx->tma_min_tiles > 0 ? (long long)ctx->tma_min_tiles
                                            : (long long)ctx->num_sms * 8))
    return -1;
  double *apz = f.apz;
  f.apz = nullptr;
  f.ncols = 0;
  TmaHostCheck chk;
  chk.nfix = F::NFIX;
  f.tstreams(chk);
  for (int j = 0; j < m; j++) chk.n(F::NFIX + j, f.V.p[j]);
  if (apz && (((uintptr_t)apz) & 15)) return -1;
  if constexpr (DOTS) chk.fixed[F::S_LIN] = chk.fixed[F::S_T] = true;
  int nslots = 0;
  plan.nmap[0] = plan.nmap[1] = plan.nmap[2] = 0ull;
  for (int i = 0; i < 24; i++) plan.noff[i] = 0u;
  for (int i = 0; i < F::NFIX; i++) {
    plan.nmap[i >> 3] |= (unsigned long long)nslots << ((i & 7) * 8);
    plan.noff[i] = (unsigned)nslots * (unsigned)(ROWS * 8);
    if (chk.fixed[i]) nslots++;
  }
  plan.col_base = nslots;
  nslots += m;
  plan.npw = PCU_TMA_NPW;
  if (!chk.aligned || chk.copies > 64 * plan.npw) return -1;
  plan.stage_bytes = nslots * ROWS * 8;
  const int scratch = PCU_WT_TEAMS * PCU_WT_TEAMW * (DOTS ? 3 : 2) * ROWS * 8;
  // 227 KB per CTA less the static shared memory (mbarriers, the reduction's staging
  // rows, the teams' dot products)
  const int budget = PCU_TMA_SMEM_BUDGET - (DOTS ? PCU_WT_TEAMS * PCU_MAX_COLS * 8 : 0);
  int fit = (budget - scratch) / plan.stage_bytes;
  if (fit > PCU_TMA_MAXSTAGES) fit = PCU_TMA_MAXSTAGES;
  if (fit < 2) return -1;
  plan.nstages = fit;
  plan.scr_off = (unsigned)(plan.nstages * plan.stage_bytes);
  plan.m = m;
  plan.nca = apz ? f.nca : 0;
  plan.reverse = (F::REVERSE && !ctx->no_reverse) ? 1 : 0;
  const size_t smem = (size_t)plan.nstages * plan.stage_bytes + scratch;
  static size_t attr_smem[PCU_MAX_DEVICES] = {0};
  const int dev = ctx->device >= 0 && ctx->device < PCU_MAX_DEVICES ? ctx->device : 0;
  if (attr_smem[dev] < smem || ctx->device >= PCU_MAX_DEVICES) {
    PCU_CUDA_OK(cudaFuncSetAttribute(wide_tile_kernel<F, DOTS, MAXJ>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[dev] = smem;
  }
  const int grid = ctx->tma_grid > 0 && ctx->tma_grid < ctx->num_sms ? ctx->tma_grid : ctx->num_sms;
  if (DOTS && ctx->big_reserve((size_t)m, (size_t)grid * m)) return 1;
  ctx->prof_begin(DOTS ? "Pass2R1W" : "Pass2SW");
  wide_tile_kernel<F, DOTS, MAXJ><<<grid, PCU_TMA_MAXWARPS * 32, smem, ctx->stream>>>(
      f, rb, plan, apz, ctx->d_big_partials, ctx->d_counter, ctx->d_big);
  ctx->prof_end();
  ctx->launches++;
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}
