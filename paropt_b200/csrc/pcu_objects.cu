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

struct pcu_blockmat {
  pcu_ctx *ctx = nullptr;
  int nvars = 0, nwcon = 0;
  WDesc wd;
  pcu_vec *Cw = nullptr;
  pcu_vec *Dinv = nullptr;  // kept from factor() like the reference (SM.cpp:44-58)
  int factored = 0;
};

struct pcu_qn {
  QuasiNewton q;
};

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
  delete m;
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
