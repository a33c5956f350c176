// pcu_objects.cu -- the two remaining abstract classes of the drop-in boundary as
// stand-alone C-ABI objects:
//   pcu_blockmat  <->  ParOptQuasiDefMat / ParOptQuasiDefBlockMat (ParOptSparseMat.h:18-104)
//   pcu_qn        <->  ParOptCompactQuasiNewton / ParOptLBFGS / ParOptLSR1
//                      (ParOptQuasiNewton.h:32-67, 76-213)
// The interior-point core fuses both into its KKT passes; these handles serve
// reference code that keeps its own ParOptInteriorPoint / ParOptTrustRegion and
// only swaps the objects behind createQuasiDefMat() / setQuasiNewton().
#include <string.h>

#include <string>

#include "pcu_ip.cuh"

#define launch_tile pcu_launch_tile
WDesc pcu_make_wdesc(const pcu_weighting &w, int nvars);

// Block form (nwblock > 1): descriptor as seen by the kernels
#define PCU_BLK_MAXNB 8
#define PCU_BLK_MAXNW 64
struct BlockDesc {
  int nblocks, nw, nb;
  long long wstart, wstride;
  double coef[PCU_BLK_MAXNB * PCU_BLK_MAXNW];  // [nb][nw]
};

struct pcu_blockmat {
  pcu_ctx *ctx = nullptr;
  int nvars = 0, nwcon = 0;
  WDesc wd;
  int nb = 1;                 // nwblock
  BlockDesc *blk = nullptr;   // nb > 1 (host copy; passed to the kernels by value)
  pcu_vec *Cw = nullptr;
  pcu_vec *Dinv = nullptr;  // kept from factor() like the reference (SM.cpp:44-58)
  int factored = 0;
};

struct pcu_qn {
  QuasiNewton q;
};

// ------------------------------------------------------ block form (nwblock > 1)
// One thread per block, the nb x nb matrix in registers (NB is a template
// parameter: fully unrolled packed-upper Cholesky, LAPACK dpptrf / dpptrs "U").
// Packed upper, column-major: entry (i, j), i <= j, at i + j (j + 1) / 2.
template <int NB>
__global__ void __launch_bounds__(128)
    block_factor_kernel(const BlockDesc d, const double *__restrict__ Dinv,
                        const double *__restrict__ Cdiag, double *__restrict__ Cw,
                        int *__restrict__ bad) {
  constexpr int NP = NB * (NB + 1) / 2;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.nblocks) return;
  double E[NP];
#pragma unroll
  for (int e = 0; e < NP; e++) E[e] = 0.0;
#pragma unroll
  for (int j = 0; j < NB; j++) E[j + j * (j + 1) / 2] = Cdiag[b * NB + j];
  const long long j0 = d.wstart + b * d.wstride;
  for (int k = 0; k < d.nw; k++) {  // Cw += Aw Dinv Aw^T (addSparseInnerProduct)
    const double dv = Dinv[j0 + k];
#pragma unroll
    for (int j = 0; j < NB; j++) {
      const double cj = d.coef[j * d.nw + k] * dv;
#pragma unroll
      for (int i = 0; i <= j; i++) E[i + j * (j + 1) / 2] += d.coef[i * d.nw + k] * cj;
    }
  }
  // dpptrf "U": A = U^T U, column by column
  int info = 0;
#pragma unroll
  for (int j = 0; j < NB; j++) {
    // U(0:j-1, j) = U(0:j-1, 0:j-1)^-T A(0:j-1, j)
#pragma unroll
    for (int i = 0; i < j; i++) {
      double v = E[i + j * (j + 1) / 2];
#pragma unroll
      for (int l = 0; l < i; l++) v -= E[l + i * (i + 1) / 2] * E[l + j * (j + 1) / 2];
      E[i + j * (j + 1) / 2] = v / E[i + i * (i + 1) / 2];
    }
    double ajj = E[j + j * (j + 1) / 2];
#pragma unroll
    for (int l = 0; l < j; l++) ajj -= E[l + j * (j + 1) / 2] * E[l + j * (j + 1) / 2];
    if (!(ajj > 0.0)) {
      if (!info) info = j + 1;
      ajj = 1.0;  // keep going with a harmless pivot; the caller sees the failure
    }
    E[j + j * (j + 1) / 2] = sqrt(ajj);
  }
  if (info) atomicMin(bad, (int)(b * NB) + info);
#pragma unroll
  for (int e = 0; e < NP; e++) Cw[b * NP + e] = E[e];
}

template <int NB>
__global__ void __launch_bounds__(128)
    block_apply_kernel(const BlockDesc d, const double *__restrict__ bx,
                       const double *__restrict__ bw, const double *__restrict__ Dinv,
                       const double *__restrict__ Cw, double *__restrict__ yx,
                       double *__restrict__ yw) {
  constexpr int NP = NB * (NB + 1) / 2;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.nblocks) return;
  const long long j0 = d.wstart + b * d.wstride;
  double r[NB], U[NP];
#pragma unroll
  for (int j = 0; j < NB; j++) r[j] = bw ? bw[b * NB + j] : 0.0;
#pragma unroll
  for (int e = 0; e < NP; e++) U[e] = Cw[b * NP + e];
  for (int k = 0; k < d.nw; k++) {  // yw = bw - Aw (Dinv bx)
    const double t = Dinv[j0 + k] * bx[j0 + k];
#pragma unroll
    for (int j = 0; j < NB; j++) r[j] -= d.coef[j * d.nw + k] * t;
  }
  // dpptrs "U": U^T z = r, then U y = z
#pragma unroll
  for (int j = 0; j < NB; j++) {
#pragma unroll
    for (int l = 0; l < j; l++) r[j] -= U[l + j * (j + 1) / 2] * r[l];
    r[j] /= U[j + j * (j + 1) / 2];
  }
#pragma unroll
  for (int j = NB - 1; j >= 0; j--) {
    r[j] /= U[j + j * (j + 1) / 2];
#pragma unroll
    for (int l = 0; l < j; l++) r[l] -= U[l + j * (j + 1) / 2] * r[j];
  }
#pragma unroll
  for (int j = 0; j < NB; j++) yw[b * NB + j] = r[j];
  for (int k = 0; k < d.nw; k++) {  // yx = Dinv (bx + Aw^T yw)
    double t = bx[j0 + k];
#pragma unroll
    for (int j = 0; j < NB; j++) t += d.coef[j * d.nw + k] * r[j];
    yx[j0 + k] = Dinv[j0 + k] * t;
  }
}

// yx = Dinv bx on the variables outside every block
__global__ void block_apply_rest_kernel(const BlockDesc d, long long n,
                                        const double *__restrict__ bx,
                                        const double *__restrict__ Dinv,
                                        double *__restrict__ yx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long end = d.wstart + (long long)d.nblocks * d.wstride;
  bool inside = false;
  if (d.nblocks > 0 && i >= d.wstart && i < end) inside = ((i - d.wstart) % d.wstride) < d.nw;
  if (!inside) yx[i] = Dinv[i] * bx[i];
}

#define PCU_BLK_DISPATCH(NBV, CALL)            \
  switch (NBV) {                               \
    case 1: { constexpr int NB = 1; CALL; } break; \
    case 2: { constexpr int NB = 2; CALL; } break; \
    case 3: { constexpr int NB = 3; CALL; } break; \
    case 4: { constexpr int NB = 4; CALL; } break; \
    case 5: { constexpr int NB = 5; CALL; } break; \
    case 6: { constexpr int NB = 6; CALL; } break; \
    case 7: { constexpr int NB = 7; CALL; } break; \
    default: { constexpr int NB = 8; CALL; } break; \
  }

static int blocks_factor(pcu_blockmat *m, pcu_vec *Dinv, pcu_vec *Cdiag) {
  pcu_ctx *ctx = m->ctx;
  const BlockDesc &d = *m->blk;
  int *bad = nullptr;
  PCU_CUDA_OK(cudaMalloc(&bad, sizeof(int)));
  const int big = 0x7fffffff;
  PCU_CUDA_OK(cudaMemcpyAsync(bad, &big, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (d.nblocks > 0) {
    const int grid = (d.nblocks + 127) / 128;
    PCU_BLK_DISPATCH(d.nb, (block_factor_kernel<NB><<<grid, 128, 0, ctx->stream>>>(
                               d, Dinv->d, Cdiag->d, m->Cw->d, bad)));
    ctx->launches++;
  }
  int host_bad = big;
  PCU_CUDA_OK(cudaMemcpyAsync(&host_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PCU_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  cudaFree(bad);
  PCU_CUDA_OK(cudaGetLastError());
  return host_bad == big ? 0 : host_bad;
}

static int blocks_apply(pcu_blockmat *m, pcu_vec *bx, pcu_vec *bw, pcu_vec *yx, pcu_vec *yw) {
  pcu_ctx *ctx = m->ctx;
  const BlockDesc &d = *m->blk;
  if (d.nblocks > 0) {
    const int grid = (d.nblocks + 127) / 128;
    PCU_BLK_DISPATCH(d.nb, (block_apply_kernel<NB><<<grid, 128, 0, ctx->stream>>>(
                               d, bx->d, bw ? bw->d : nullptr, m->Dinv->d, m->Cw->d, yx->d,
                               yw->d)));
    ctx->launches++;
  }
  if (m->nvars > 0) {
    block_apply_rest_kernel<<<(m->nvars + 255) / 256, 256, 0, ctx->stream>>>(
        d, m->nvars, bx->d, m->Dinv->d, yx->d);
    ctx->launches++;
  }
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" {

// ------------------------------------------------------------------ blockmat
pcu_blockmat *pcu_blockmat_create(pcu_ctx *ctx, int nvars, const pcu_weighting *w) {
  if (!ctx || nvars < 0) return nullptr;
  pcu_blockmat *m = new pcu_blockmat;
  m->ctx = ctx;
  m->nvars = nvars;
  pcu_weighting none;
  memset(&none, 0, sizeof(none));
  const pcu_weighting &ww = w ? *w : none;
  if (pcu_validate_weighting(&ww, nvars, 0, 0, ww.nwcon, "pcu_blockmat_create")) {
    delete m;
    return nullptr;
  }
  m->nwcon = ww.nwcon;
  m->wd = pcu_make_wdesc(ww, nvars);
  m->Cw = pcu_vec_create(ctx, m->nwcon);
  if (!m->Cw) {
    delete m;
    return nullptr;
  }
  return m;
}

void pcu_blockmat_destroy(pcu_blockmat *m) {
  if (!m) return;
  pcu_vec_destroy(m->Cw);
  delete m->blk;
  delete m;
}

pcu_blockmat *pcu_blockmat_create_blocks(pcu_ctx *ctx, int nvars,
                                         const pcu_block_weighting *b) {
  if (!ctx || !b || nvars < 0) return nullptr;
  const char *why = nullptr;
  if (b->nb < 1 || b->nb > PCU_BLK_MAXNB) why = "nb outside 1..8";
  else if (b->nw < 1 || b->nw > PCU_BLK_MAXNW) why = "nw outside 1..64";
  else if (b->nblocks < 0) why = "negative number of blocks";
  else if (!b->coef) why = "no coefficient matrix";
  else if (b->nblocks > 0 &&
           (b->wstart < 0 || b->wstride < b->nw ||
            (long long)b->wstart + (long long)(b->nblocks - 1) * b->wstride + b->nw > (long long)nvars))
    why = "blocks overlap or reach outside the vector";
  if (why) {
    fprintf(stderr, "paropt_b200: pcu_blockmat_create_blocks: %s\n", why);
    return nullptr;
  }
  pcu_blockmat *m = new pcu_blockmat;
  m->ctx = ctx;
  m->nvars = nvars;
  m->nb = b->nb;
  m->nwcon = b->nblocks * b->nb;
  memset(&m->wd, 0, sizeof(m->wd));
  m->blk = new BlockDesc;
  memset(m->blk, 0, sizeof(BlockDesc));
  m->blk->nblocks = b->nblocks;
  m->blk->nw = b->nw;
  m->blk->nb = b->nb;
  m->blk->wstart = b->wstart;
  m->blk->wstride = b->wstride;
  memcpy(m->blk->coef, b->coef, sizeof(double) * (size_t)b->nb * b->nw);
  // packed upper triangle per block (the reference's Cw layout, SM.cpp:24)
  m->Cw = pcu_vec_create(ctx, b->nblocks * (b->nb * (b->nb + 1) / 2));
  if (!m->Cw) {
    delete m->blk;
    delete m;
    return nullptr;
  }
  return m;
}

// int ParOptQuasiDefMat::factor(x, Dinv, Cdiag): 0 = ok, else 1 + the first
// failing row (global on this rank; the reference returns the row too).  x is
// accepted for signature parity: the weighting Jacobian does not depend on it.
int pcu_blockmat_factor(pcu_blockmat *m, pcu_vec *x, pcu_vec *Dinv, pcu_vec *Cdiag) {
  (void)x;
  if (!m || !Dinv || !Cdiag || Dinv->n != m->nvars || Cdiag->n != m->nwcon) return -1;
  pcu_vec_ready(Dinv);
  pcu_vec_ready(Cdiag);
  m->Dinv = Dinv;
  m->factored = 1;
  if (m->blk) return blocks_factor(m, Dinv, Cdiag);
  BlockFactorF f;
  f.Dinv = Dinv->d;
  f.Cdiag = Cdiag->d;
  f.Cw = m->Cw->d;
  RedBuf rb = m->ctx->redbuf(0, 1, 0);
  if (launch_tile(m->ctx, f, m->nvars, m->wd, rb)) return -1;
  double bad = 0.0;
  if (m->ctx->fetch(&bad)) return -1;
  return (int)bad;
}

static int blockmat_apply(pcu_blockmat *m, pcu_vec *bx, pcu_vec *bw, pcu_vec *yx,
                          pcu_vec *yw) {
  if (!m || !m->factored || !bx || !yx || !yw || bx->n != m->nvars || yx->n != m->nvars ||
      yw->n != m->nwcon || (bw && bw->n != m->nwcon))
    return 1;
  pcu_vec_ready(bx);
  pcu_vec_ready(bw);
  pcu_vec_ready(yx);
  pcu_vec_ready(yw);
  pcu_vec_ready(m->Dinv);
  if (m->blk) return blocks_apply(m, bx, bw, yx, yw);
  BlockApplyF f;
  f.bx = bx->d;
  f.bw = bw ? bw->d : nullptr;
  f.Dinv = m->Dinv->d;
  f.Cw = m->Cw->d;
  f.yx = yx->d;
  f.yw = yw->d;
  const RedBuf none = {nullptr, nullptr, nullptr, 0};
  return launch_tile(m->ctx, f, m->nvars, m->wd, none);
}
int pcu_blockmat_apply3(pcu_blockmat *m, pcu_vec *bx, pcu_vec *yx, pcu_vec *yw) {
  return blockmat_apply(m, bx, nullptr, yx, yw);
}
int pcu_blockmat_apply4(pcu_blockmat *m, pcu_vec *bx, pcu_vec *bw, pcu_vec *yx,
                        pcu_vec *yw) {
  return blockmat_apply(m, bx, bw, yx, yw);
}

// ------------------------------------------------------------------------ qn
pcu_qn *pcu_qn_create(pcu_ctx *ctx, int nvars, const char *qn_type, int subspace) {
  if (!ctx || !qn_type || subspace < 1 || nvars < 0) return nullptr;
  const std::string t(qn_type);
  int kind;
  if (t == "bfgs") kind = 0;
  else if (t == "sr1") kind = 1;
  else {
    fprintf(stderr, "paropt_b200: pcu_qn_create: unknown qn_type %s\n", qn_type);
    return nullptr;
  }
  pcu_qn *h = new pcu_qn;
  if (h->q.init(ctx, nvars, kind, subspace)) {
    delete h;
    return nullptr;
  }
  return h;
}
void pcu_qn_destroy(pcu_qn *h) { delete h; }

// setQuasiNewtonUpdateType / setInitDiagonalType (QN.h:100-107, 58)
int pcu_qn_set_option(pcu_qn *h, const char *name, const char *value) {
  if (!h || !name || !value) return 1;
  const std::string k(name), v(value);
  if (k == "qn_update_type") {
    if (v == "skip_negative_curvature") h->q.damped = 0;
    else if (v == "damped_update") h->q.damped = 1;
    else return 1;
  } else if (k == "qn_diag_type") {
    // the inner_* values are accepted and act as yty_over_yts, exactly as in the
    // reference (QN.cpp:200-204 tests for PAROPT_YTS_OVER_STS only)
    if (v == "yts_over_sts") h->q.diag_yts_over_sts = 1;
    else if (v == "yty_over_yts" || v == "inner_yty_over_yts" || v == "inner_yts_over_sts")
      h->q.diag_yts_over_sts = 0;
    else return 1;
  } else {
    return 1;
  }
  return 0;
}
int pcu_qn_reset(pcu_qn *h) {
  if (!h) return 1;
  h->q.reset();
  return 0;
}
// getMaxLimitedMemorySize (QN.cpp:127, 603): 2 m for L-BFGS, m for L-SR1
int pcu_qn_max_size(pcu_qn *h) { return h ? h->q.max_size() : 0; }

// int update(x, z, zw, s, y): 0 normal, 1 damped, 2 skipped (QN.cpp:162-334, 636-747)
int pcu_qn_update(pcu_qn *h, pcu_vec *s, pcu_vec *y, int *update_type) {
  if (!h || !s || !y) return 1;
  pcu_vec_ready(s);
  pcu_vec_ready(y);
  double yy, ys, ss;
  if (pcu_vec_dot(y, y, &yy) || pcu_vec_dot(y, s, &ys) || pcu_vec_dot(s, s, &ss)) return 1;
  int ut = 0;
  if (h->q.update(s, y, yy, ys, ss, nullptr, &ut)) return 1;
  if (update_type) *update_type = ut;
  return 0;
}
// y = B x (QN.cpp:390-418, 760-778)
int pcu_qn_mult(pcu_qn *h, pcu_vec *x, pcu_vec *y) {
  if (!h || !x || !y) return 1;
  pcu_vec_ready(x);
  pcu_vec_ready(y);
  return h->q.mult(x, y);
}
// y += alpha B x (QN.cpp:432-459, 792-809)
int pcu_qn_mult_add(pcu_qn *h, double alpha, pcu_vec *x, pcu_vec *y) {
  if (!h || !x || !y) return 1;
  pcu_vec_ready(x);
  pcu_vec_ready(y);
  if (h->q.mult(x, h->q.r)) return 1;
  return pcu_vec_axpy(y, alpha, h->q.r);
}
// int getCompactMat(&b0, &d, &M, &Z): returns the width q; d0[q], M[q*q]
// column-major, Z[q] borrowed handles (QN.cpp:471-487, 821-837)
int pcu_qn_compact(pcu_qn *h, double *b0, double *d0, double *M, pcu_vec **Z) {
  if (!h) return -1;
  const int q = h->q.size();
  if (b0) *b0 = h->q.b0;
  if (d0 && q > 0) memcpy(d0, h->q.d0.data(), sizeof(double) * q);
  if (M && q > 0) memcpy(M, h->q.M.data(), sizeof(double) * q * q);
  if (Z) {
    const int ms = h->q.msub;
    for (int i = 0; i < q; i++) {
      if (h->q.type == 0) Z[i] = i < ms ? h->q.S[i] : h->q.Y[i - ms];
      else Z[i] = h->q.Zs[i];
    }
  }
  return q;
}

// ParOptInteriorPoint::setQuasiNewton (IP.cpp:1193-1235): the optimizer uses -- and,
// with use_quasi_newton_update, updates -- the caller's object; NULL restores the
// optimizer's own (options qn_type / qn_subspace_size).  The handle must outlive
// its use by the optimizer.
int pcu_ip_set_quasi_newton(pcu_ip *ip, pcu_qn *h) {
  if (!ip) return 1;
  if (!ip->qn_external) delete ip->qn;
  ip->qn = nullptr;
  ip->qn_external = 0;
  ip->qn_built_size = -1;
  if (h) {
    if (h->q.n != ip->nvars || ip->ncon + h->q.max_size() + 1 > PCU_MAX_COLS) {
      fprintf(stderr, "paropt_b200: pcu_ip_set_quasi_newton: size mismatch\n");
      return 1;
    }
    ip->qn = &h->q;
    ip->qn_external = 1;
  }
  return 0;
}

}  // extern "C"
