/*
  oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the
  product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
  --impl reference legs execute the binary built from it).

  Runs the UNMODIFIED reference ParOptInteriorPoint (sources compiled where they
  lie under /root/reference/src, see oracle/Makefile) on the synthetic problems
  of DESIGN.md section "Synthetic problems" and records a full-precision
  per-iteration history.  The reference's text log only carries 2 significant
  digits (ParOptInteriorPoint.cpp:4777-4801), so the history is taken through
  the reference's own writeOutput hook (ParOptInteriorPoint.cpp:4620-4631,
  write_output_frequency=1): at the top of every major iteration the hook
  re-evaluates computeComp / computeKKTRes / computeResNorm (the reference's own
  private methods; this translation unit sees the class with `private` spelled
  `public`, the reference objects themselves are compiled untouched) and prints
  one JSON line.

  Problems (this file restates the generator spec; it shares no code with the
  CUDA library or the numpy oracle so that the three can check each other):
    problem=sepquad     separable/Householder convex QP, dense linear
                        constraints, optional multi-material weighting
                        constraints (configs C2..C5 of BASELINE.json)
    problem=rosenbrock  the problem class of
                        /root/reference/examples/rosenbrock/rosenbrock.cpp:9-199
                        (config C1), driven through ParOptInteriorPoint directly
*/
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <complex>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "mpi.h"

// Every standard header the reference headers pull in is included above, so
// the access-specifier override below only touches the reference's classes.
#define private public
#include "ParOptInteriorPoint.h"
#include "ParOptOptimizer.h"
#undef private

// -DPCU_ADAPTERS (oracle/Makefile target _ref/adapter_driver): the same driver, the
// same UNMODIFIED reference objects, but the problem classes hand out CUDA vectors,
// the CUDA block matrix and the CUDA compact quasi-Newton object through the
// reference's own factories (src/ParOptProblem.h:58,65,72; setQuasiNewton,
// IP.cpp:1193) -- tests/adapters/paropt_cuda_adapters.h over include/paropt_b200.h.
#ifdef PCU_ADAPTERS
#include "paropt_cuda_adapters.h"
#endif

// ---------------------------------------------------------------------------
// Counter-based generator ("Synthetic problems")
// ---------------------------------------------------------------------------
static inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline uint64_t stream_key(uint64_t seed, uint64_t stream) {
  return splitmix64(seed ^ (stream * 0x9E3779B97F4A7C15ULL));
}
static inline double uniform01(uint64_t key, uint64_t idx) {
  return (double)(splitmix64(key + idx) >> 11) * (1.0 / 9007199254740992.0);
}

struct SepQuadParams {
  long ntotal = 1000;
  int ncon = 1;
  int nw = 0;  // variables per weighting block; 0 = no weighting constraints
  int nb = 1;  // sparse constraints per block (nwblock of ParOptQuasiDefBlockMat): 1 or 2
  uint64_t seed = 0;
  double lam_min = 1.0, lam_max = 1e3;
  double b_lo = 0.0, b_w = 1.0;
  double a_lo = 0.0, a_w = 1.0;
  double beta_c = 0.0, beta_n = 0.0, beta_u = 1.0;
  double x0_lo[2] = {-2.0, -2.0}, x0_w[2] = {1.0, 1.0};
  double lb[2] = {-5.0, -5.0}, ub[2] = {5.0, 5.0};
  int householder = 0;
};

class HistoryProblem : public ParOptProblem {
 public:
  HistoryProblem(MPI_Comm comm) : ParOptProblem(comm) {
    ip = NULL;
    hist = NULL;
    have_prev = 0;
    callback_time = 0.0;
    tr_mode = 0;
  }
  void writeOutput(int iter, ParOptVec *xvec);
#ifdef PCU_ADAPTERS
  ParOptVec *createDesignVec() {
    int nv, nc, nw;
    getProblemSizes(&nv, &nc, &nw);
    return new ParOptCudaVec(ParOptCudaContext(), nv);
  }
  ParOptVec *createConstraintVec() {
    int nv, nc, nw;
    getProblemSizes(&nv, &nc, &nw);
    return new ParOptCudaVec(ParOptCudaContext(), nw);
  }
#endif

  ParOptInteriorPoint *ip;
  FILE *hist;
  std::vector<double> xprev;
  int have_prev;
  double callback_time;
  int tr_mode;  // algorithm=tr: writeOutput is the trust-region hook (ParOptTrustRegion.cpp:1559)
};

static double vsum(ParOptVec *v) {
  double *a;
  int n = v->getArray(&a);
  double s = 0.0;
  for (int i = 0; i < n; i++) s += a[i];
  double out = 0.0;
  MPI_Allreduce(&s, &out, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
  return out;
}

static void print_arr(FILE *fp, const char *name, const double *a, int n) {
  fprintf(fp, "\"%s\": [", name);
  for (int i = 0; i < n; i++) fprintf(fp, "%s%.17g", i ? ", " : "", a[i]);
  fprintf(fp, "]");
}

void HistoryProblem::writeOutput(int iter, ParOptVec *xvec) {
  int rank;
  MPI_Comm_rank(comm, &rank);
  if (tr_mode) {
    // one record per trust-region iteration: the centre point x_k at full precision
    // (checksums + the first entries); the scalars of the iteration are in the
    // reference's own trust-region log (tr_output_file), parsed by make_golden.py
    double *x;
    int n = xvec->getArray(&x);
    double xs = vsum(xvec), xn = xvec->norm(), xm = xvec->maxabs();
    if (rank == 0 && hist) {
      fprintf(hist, "{\"tr_iter\": %d, \"xsum\": %.17g, \"xnorm\": %.17g, \"xmaxabs\": %.17g, ",
              iter, xs, xn, xm);
      print_arr(hist, "xhead", x, n < 8 ? n : 8);
      fprintf(hist, "}\n");
      fflush(hist);
    }
    return;
  }
  if (!ip) return;
  ParOptInteriorPoint::ParOptVars &v = ip->variables;

  // Same calls, same order as the top of the major loop
  // (ParOptInteriorPoint.cpp:4656, 4670-4671)
  double comp = ip->computeComp(v);
  double max_prime, max_dual, max_infeas, res_norm;
  ip->computeKKTRes(v, ip->barrier_param, ip->residual);
  // in the norm the optimizer itself uses (option norm_type, IP.cpp:4462-4470)
  ParOptNormType ntype = PAROPT_INFTY_NORM;
  const char *nname = ip->options->getEnumOption("norm_type");
  if (strcmp(nname, "l1") == 0) ntype = PAROPT_L1_NORM;
  else if (strcmp(nname, "l2") == 0) ntype = PAROPT_L2_NORM;
  ip->computeResNorm(ntype, ip->residual, &max_prime, &max_dual, &max_infeas,
                     &res_norm);

  // Recover the accepted line-search step length from x_k - x_{k-1} and the
  // (already alpha_x-scaled) step still stored in ip->update.x
  double alpha = 0.0, pnorm2 = 0.0;
  double *x, *px;
  int n = xvec->getArray(&x);
  ip->update.x->getArray(&px);
  if (have_prev) {
    double num = 0.0, den = 0.0;
    for (int i = 0; i < n; i++) {
      num += (x[i] - xprev[i]) * px[i];
      den += px[i] * px[i];
    }
    double in[2] = {num, den}, out[2];
    MPI_Allreduce(in, out, 2, MPI_DOUBLE, MPI_SUM, comm);
    pnorm2 = out[1];
    alpha = (out[1] > 0.0) ? out[0] / out[1] : 0.0;
  }
  xprev.assign(x, x + n);
  have_prev = 1;

  double b0 = 0.0;
  int qsize = 0;
  if (ip->qn) qsize = ip->qn->getCompactMat(&b0, NULL, NULL, NULL);

  double sums[9] = {vsum(v.x),   vsum(v.zl), vsum(v.zu),  vsum(v.zw), vsum(v.sw),
                    vsum(v.tw),  vsum(v.zsw), vsum(v.ztw), 0.0};
  double xnorm = v.x->norm();
  double zlnorm = v.zl->norm(), zunorm = v.zu->norm();
  double zwnorm = v.zw->norm();
  double gmax = ip->g->maxabs();

  if (rank == 0 && hist) {
    int ncon = ip->ncon;
    fprintf(hist,
            "{\"iter\": %d, \"wall\": %.6f, \"fobj\": %.17g, \"mu\": %.17g, \"rho\": %.17g, "
            "\"comp\": %.17g, \"max_prime\": %.17g, \"max_dual\": %.17g, "
            "\"max_infeas\": %.17g, \"res_norm\": %.17g, \"neval\": %d, "
            "\"ngeval\": %d, \"nhvec\": %d, \"alpha\": %.17g, \"pnorm2\": %.17g, "
            "\"qn_b0\": %.17g, \"qn_size\": %d, \"xsum\": %.17g, "
            "\"xnorm\": %.17g, \"zlsum\": %.17g, \"zusum\": %.17g, "
            "\"zlnorm\": %.17g, \"zunorm\": %.17g, \"zwsum\": %.17g, "
            "\"zwnorm\": %.17g, \"swsum\": %.17g, \"twsum\": %.17g, "
            "\"zswsum\": %.17g, \"ztwsum\": %.17g, \"gmax\": %.17g, ",
            iter, MPI_Wtime(), ip->fobj, ip->barrier_param, ip->rho_penalty_search, comp,
            max_prime, max_dual, max_infeas, res_norm, ip->neval, ip->ngeval, ip->nhvec,
            alpha, pnorm2, b0, qsize, sums[0], xnorm, sums[1], sums[2], zlnorm,
            zunorm, sums[3], zwnorm, sums[4], sums[5], sums[6], sums[7], gmax);
    print_arr(hist, "c", ip->c, ncon);
    fprintf(hist, ", ");
    print_arr(hist, "z", v.z, ncon);
    fprintf(hist, ", ");
    print_arr(hist, "s", v.s, ncon);
    fprintf(hist, ", ");
    print_arr(hist, "t", v.t, ncon);
    fprintf(hist, ", ");
    print_arr(hist, "zs", v.zs, ncon);
    fprintf(hist, ", ");
    print_arr(hist, "zt", v.zt, ncon);
    fprintf(hist, "}\n");
    fflush(hist);
  }
}

// ---------------------------------------------------------------------------
// sepquad
// ---------------------------------------------------------------------------
class SepQuad : public HistoryProblem {
 public:
  SepQuad(MPI_Comm comm, const SepQuadParams &params) : HistoryProblem(comm) {
    p = params;
    int rank, size;
    MPI_Comm_rank(comm, &rank);
    MPI_Comm_size(comm, &size);
    // Block-row partition in units of one weighting block (or 1 variable)
    long unit = p.nw > 0 ? p.nw : 1;
    long nunits = p.ntotal / unit;
    long u0 = (nunits * rank) / size, u1 = (nunits * (rank + 1)) / size;
    offset = u0 * unit;
    n = (int)((u1 - u0) * unit);
    if (rank == size - 1) n = (int)(p.ntotal - offset);
    nblk = p.nw > 0 ? (int)(u1 - u0) : 0;
    nwc = nblk * p.nb;
    // second row of a block (nb = 2): cw = 1.5 - sum_k ((k % 3) + 1) / 3 x_k
    c1.resize(p.nw > 0 ? p.nw : 1);
    for (int k = 0; k < p.nw; k++) c1[k] = -((k % 3) + 1) / 3.0;
    setProblemSizes(n, p.ncon, nwc);
    setNumInequalities(p.ncon, nwc);

    lam.resize(n);
    b.resize(n);
    vh.resize(n);
    uint64_t klam = stream_key(p.seed, 1), kb = stream_key(p.seed, 2);
    uint64_t kv = stream_key(p.seed, 7);
    double loc = 0.0;
    for (int i = 0; i < n; i++) {
      uint64_t gi = (uint64_t)(offset + i);
      lam[i] = p.lam_min + (p.lam_max - p.lam_min) * uniform01(klam, gi);
      b[i] = p.b_lo + p.b_w * uniform01(kb, gi);
      vh[i] = 0.5 + uniform01(kv, gi);
      loc += vh[i] * vh[i];
    }
    vtv = 0.0;
    MPI_Allreduce(&loc, &vtv, 1, MPI_DOUBLE, MPI_SUM, comm);
    beta.resize(p.ncon);
    uint64_t kbeta = stream_key(p.seed, 5);
    for (int j = 0; j < p.ncon; j++) {
      beta[j] = p.beta_c + p.beta_n * (double)p.ntotal +
                p.beta_u * uniform01(kbeta, (uint64_t)j);
    }
    ytmp.resize(n);
  }

  ParOptQuasiDefMat *createQuasiDefMat() {
#ifdef PCU_ADAPTERS
    if (p.nb > 1) {
      // block form: nb rows per block, dense nb x nw coefficient matrix
      std::vector<double> coef((size_t)p.nb * p.nw);
      for (int k = 0; k < p.nw; k++) {
        coef[k] = k == 0 ? 1.0 : -1.0;
        coef[p.nw + k] = c1[k];
      }
      pcu_block_weighting bw;
      memset(&bw, 0, sizeof(bw));
      bw.nblocks = nblk;
      bw.wstart = 0;
      bw.nw = p.nw;
      bw.wstride = p.nw;
      bw.nb = p.nb;
      bw.coef = coef.data();
      return new ParOptCudaQuasiDefBlockMat(ParOptCudaContext(), n, &bw);
    }
    pcu_weighting w;
    memset(&w, 0, sizeof(w));
    if (nwc > 0) {
      w.nwcon = nwc;
      w.wstart = 0;
      w.nw = p.nw;
      w.wstride = p.nw;
      w.coef0 = 1.0;
      w.coef_rest = -1.0;
    }
    return new ParOptCudaQuasiDefBlockMat(ParOptCudaContext(), n, &w);
#else
    int nwblock = (nwc > 0) ? p.nb : 0;
    return new ParOptQuasiDefBlockMat(this, nwblock);
#endif
  }

  int cls(int i) const { return (p.nw > 0 && (i % p.nw) != 0) ? 1 : 0; }
  double acoef(int j, uint64_t gi) const {
    return p.a_lo + p.a_w * uniform01(stream_key(p.seed, 100 + (uint64_t)j), gi);
  }

  void getVarsAndBounds(ParOptVec *xvec, ParOptVec *lbvec, ParOptVec *ubvec) {
    double *x, *lb, *ub;
    xvec->getArray(&x);
    lbvec->getArray(&lb);
    ubvec->getArray(&ub);
    uint64_t kx = stream_key(p.seed, 3);
    for (int i = 0; i < n; i++) {
      int k = cls(i);
      x[i] = p.x0_lo[k] + p.x0_w[k] * uniform01(kx, (uint64_t)(offset + i));
      lb[i] = p.lb[k];
      ub[i] = p.ub[k];
    }
  }

  // y = P x with P = I - 2 v v^T / (v^T v)   (householder) or y = x
  void applyP(const double *x, double *y) {
    if (!p.householder) {
      memcpy(y, x, n * sizeof(double));
      return;
    }
    double loc = 0.0, vx = 0.0;
    for (int i = 0; i < n; i++) loc += vh[i] * x[i];
    MPI_Allreduce(&loc, &vx, 1, MPI_DOUBLE, MPI_SUM, comm);
    double f = 2.0 * vx / vtv;
    for (int i = 0; i < n; i++) y[i] = x[i] - f * vh[i];
  }

  int evalObjCon(ParOptVec *xvec, ParOptScalar *fobj, ParOptScalar *cons) {
    double t0 = MPI_Wtime();
    double *x;
    xvec->getArray(&x);
    applyP(x, ytmp.data());
    std::vector<double> loc(1 + p.ncon, 0.0), out(1 + p.ncon, 0.0);
    for (int i = 0; i < n; i++) {
      loc[0] += 0.5 * lam[i] * ytmp[i] * ytmp[i] + b[i] * x[i];
    }
    for (int j = 0; j < p.ncon; j++) {
      uint64_t key = stream_key(p.seed, 100 + (uint64_t)j);
      double s = 0.0;
      for (int i = 0; i < n; i++) {
        s += (p.a_lo + p.a_w * uniform01(key, (uint64_t)(offset + i))) * x[i];
      }
      loc[1 + j] = s;
    }
    MPI_Allreduce(loc.data(), out.data(), 1 + p.ncon, MPI_DOUBLE, MPI_SUM,
                  comm);
    *fobj = out[0];
    for (int j = 0; j < p.ncon; j++) cons[j] = beta[j] + out[1 + j];
    callback_time += MPI_Wtime() - t0;
    return 0;
  }

  int evalObjConGradient(ParOptVec *xvec, ParOptVec *gvec, ParOptVec **Ac) {
    double t0 = MPI_Wtime();
    double *x, *g;
    xvec->getArray(&x);
    gvec->getArray(&g);
    applyP(x, ytmp.data());
    for (int i = 0; i < n; i++) ytmp[i] *= lam[i];
    if (p.householder) {
      double loc = 0.0, vw = 0.0;
      for (int i = 0; i < n; i++) loc += vh[i] * ytmp[i];
      MPI_Allreduce(&loc, &vw, 1, MPI_DOUBLE, MPI_SUM, comm);
      double f = 2.0 * vw / vtv;
      for (int i = 0; i < n; i++) g[i] = (ytmp[i] - f * vh[i]) + b[i];
    } else {
      for (int i = 0; i < n; i++) g[i] = ytmp[i] + b[i];
    }
    for (int j = 0; j < p.ncon; j++) {
      double *a;
      Ac[j]->getArray(&a);
      uint64_t key = stream_key(p.seed, 100 + (uint64_t)j);
      for (int i = 0; i < n; i++) {
        a[i] = p.a_lo + p.a_w * uniform01(key, (uint64_t)(offset + i));
      }
    }
    callback_time += MPI_Wtime() - t0;
    return 0;
  }

  // H = P diag(lam) P is constant (quadratic objective, linear constraints):
  // the callback of the inexact-Newton GMRES path (IP.cpp:5973)
  int evalHvecProduct(ParOptVec *, ParOptScalar *, ParOptVec *, ParOptVec *pxvec,
                      ParOptVec *hvec) {
    double *px, *h;
    pxvec->getArray(&px);
    hvec->getArray(&h);
    applyP(px, ytmp.data());
    for (int i = 0; i < n; i++) ytmp[i] *= lam[i];
    if (p.householder) {
      double loc = 0.0, vw = 0.0;
      for (int i = 0; i < n; i++) loc += vh[i] * ytmp[i];
      MPI_Allreduce(&loc, &vw, 1, MPI_DOUBLE, MPI_SUM, comm);
      double f = 2.0 * vw / vtv;
      for (int i = 0; i < n; i++) h[i] = ytmp[i] - f * vh[i];
    } else {
      for (int i = 0; i < n; i++) h[i] = ytmp[i];
    }
    return 0;
  }

  // cw_i = x[nw i] - sum_{k=1}^{nw-1} x[nw i + k]
  // (weighting form of examples/dmo_truss/dmo_truss_analysis.py:650-679); with nb = 2
  // every block carries a second row 1.5 + sum_k c1_k x[nw i + k] and the blocks of
  // Aw D^-1 Aw^T are dense 2 x 2 (packed upper: (0,0), (0,1), (1,1))
  void evalSparseCon(ParOptVec *xvec, ParOptVec *out) {
    double *x, *o;
    xvec->getArray(&x);
    out->getArray(&o);
    for (int i = 0; i < nblk; i++) {
      double s = x[p.nw * i];
      for (int k = 1; k < p.nw; k++) s -= x[p.nw * i + k];
      o[p.nb * i] = s;
      if (p.nb > 1) {
        double s1 = 1.5;
        for (int k = 0; k < p.nw; k++) s1 += c1[k] * x[p.nw * i + k];
        o[p.nb * i + 1] = s1;
      }
    }
  }
  void addSparseJacobian(ParOptScalar alpha, ParOptVec *, ParOptVec *px,
                         ParOptVec *out) {
    double *v, *o;
    px->getArray(&v);
    out->getArray(&o);
    for (int i = 0; i < nblk; i++) {
      double s = v[p.nw * i];
      for (int k = 1; k < p.nw; k++) s -= v[p.nw * i + k];
      o[p.nb * i] += alpha * s;
      if (p.nb > 1) {
        double s1 = 0.0;
        for (int k = 0; k < p.nw; k++) s1 += c1[k] * v[p.nw * i + k];
        o[p.nb * i + 1] += alpha * s1;
      }
    }
  }
  void addSparseJacobianTranspose(ParOptScalar alpha, ParOptVec *,
                                  ParOptVec *pzw, ParOptVec *out) {
    double *z, *o;
    pzw->getArray(&z);
    out->getArray(&o);
    for (int i = 0; i < nblk; i++) {
      o[p.nw * i] += alpha * z[p.nb * i];
      for (int k = 1; k < p.nw; k++) o[p.nw * i + k] -= alpha * z[p.nb * i];
      if (p.nb > 1) {
        for (int k = 0; k < p.nw; k++) o[p.nw * i + k] += alpha * c1[k] * z[p.nb * i + 1];
      }
    }
  }
  void addSparseInnerProduct(ParOptScalar alpha, ParOptVec *, ParOptVec *cvec,
                             ParOptScalar *A) {
    double *cv;
    cvec->getArray(&cv);
    if (p.nb == 1) {
      for (int i = 0; i < nblk; i++) {
        double s = 0.0;
        for (int k = 0; k < p.nw; k++) s += cv[p.nw * i + k];
        A[i] += alpha * s;
      }
    } else {
      for (int i = 0; i < nblk; i++) {
        double s00 = 0.0, s01 = 0.0, s11 = 0.0;
        for (int k = 0; k < p.nw; k++) {
          const double c0 = k == 0 ? 1.0 : -1.0, d = cv[p.nw * i + k];
          s00 += c0 * c0 * d;
          s01 += c0 * c1[k] * d;
          s11 += c1[k] * c1[k] * d;
        }
        A[3 * i] += alpha * s00;
        A[3 * i + 1] += alpha * s01;
        A[3 * i + 2] += alpha * s11;
      }
    }
  }

  SepQuadParams p;
  long offset;
  int n, nwc, nblk;
  std::vector<double> c1;
  double vtv;
  std::vector<double> lam, b, vh, beta, ytmp;
};

// ---------------------------------------------------------------------------
// sparsequad: a ParOptSparseProblem (ParOptProblem.h:301-407) -- general CSR sparse
// constraints, ParOptQuasiDefSparseMat + the reference's sparse Cholesky.
//   f = sum 1/2 lam_i x_i^2 + b_i x_i,   c_j = beta_j + a_j . x >= 0 (dense, as sepquad),
//   cw_i = rad - sum_k w_ik (x_{j(i,k)} - xc)^2 >= 0,   i < W = (n - 1) / 3,
//   j(i, .) = 3i, 3i+1, 3i+2, 3i+3 (neighbouring rows share a variable: K = C + A D^-1 A^T
//   is tridiagonal) and, for i % 5 == 0, the long-range column (7i + 11) mod n (fill).
// The Jacobian values -2 w_ik (x - xc) change with x: every factorisation sees new data.
// Single rank (the reference's sparse path is serial per rank).
// ---------------------------------------------------------------------------
class Recorder : public HistoryProblem {  // the history hook without a problem of its own
 public:
  Recorder(MPI_Comm comm) : HistoryProblem(comm) {}
  ParOptQuasiDefMat *createQuasiDefMat() { return NULL; }
  void getVarsAndBounds(ParOptVec *, ParOptVec *, ParOptVec *) {}
  int evalObjCon(ParOptVec *, ParOptScalar *, ParOptScalar *) { return 1; }
  int evalObjConGradient(ParOptVec *, ParOptVec *, ParOptVec **) { return 1; }
};

class SparseQuad : public ParOptSparseProblem {
 public:
  SparseQuad(MPI_Comm comm, const SepQuadParams &params) : ParOptSparseProblem(comm) {
    p = params;
    rec = new Recorder(comm);
    rec->incref();
    n = (int)p.ntotal;
    W = (n - 1) / 3;
    setProblemSizes(n, p.ncon, W);
    setNumInequalities(p.ncon, W);
    std::vector<int> rowp(W + 1, 0), cols;
    for (int i = 0; i < W; i++) {
      for (int k = 0; k < 4; k++) cols.push_back(3 * i + k);
      if (i % 5 == 0) {
        const int far = (int)((7L * i + 11) % n);
        if (far < 3 * i || far > 3 * i + 3) cols.push_back(far);
      }
      rowp[i + 1] = (int)cols.size();
    }
    setSparseJacobianData(rowp.data(), cols.data());
    wk.resize(cols.size());
    uint64_t kw = stream_key(p.seed, 9);
    for (size_t e = 0; e < cols.size(); e++) wk[e] = 0.5 + uniform01(kw, (uint64_t)e);
    lam.resize(n);
    b.resize(n);
    uint64_t klam = stream_key(p.seed, 1), kb = stream_key(p.seed, 2);
    for (int i = 0; i < n; i++) {
      lam[i] = p.lam_min + (p.lam_max - p.lam_min) * uniform01(klam, (uint64_t)i);
      b[i] = p.b_lo + p.b_w * uniform01(kb, (uint64_t)i);
    }
    beta.resize(p.ncon);
    uint64_t kbeta = stream_key(p.seed, 5);
    for (int j = 0; j < p.ncon; j++)
      beta[j] = p.beta_c + p.beta_n * (double)p.ntotal + p.beta_u * uniform01(kbeta, (uint64_t)j);
#ifdef PCU_ADAPTERS
    cmat = NULL;
#endif
  }
  ~SparseQuad() { rec->decref(); }
  static double rad() { return 1.0; }
  static double xc() { return 0.3; }

#ifdef PCU_ADAPTERS
  ParOptVec *createDesignVec() { return new ParOptCudaVec(ParOptCudaContext(), n); }
  ParOptVec *createConstraintVec() { return new ParOptCudaVec(ParOptCudaContext(), W); }
  ParOptQuasiDefMat *createQuasiDefMat() {
    cmat = new ParOptCudaQuasiDefSparseMat(ParOptCudaContext(), this);
    return cmat;
  }
  // the two CSR products on the device (the values are the ones of the last factor():
  // the optimizer factors after every gradient evaluation and before any product)
  void addSparseJacobian(ParOptScalar alpha, ParOptVec *x, ParOptVec *px, ParOptVec *out) {
    if (!cmat) return ParOptSparseProblem::addSparseJacobian(alpha, x, px, out);
    sync_data();
    PCU_ADAPTER_CHECK(pcu_sparsemat_mult_add(cmat->mat, alpha, ParOptCudaVec::handle(px),
                                             ParOptCudaVec::handle(out)));
  }
  void addSparseJacobianTranspose(ParOptScalar alpha, ParOptVec *x, ParOptVec *pzw,
                                  ParOptVec *out) {
    if (!cmat) return ParOptSparseProblem::addSparseJacobianTranspose(alpha, x, pzw, out);
    sync_data();
    PCU_ADAPTER_CHECK(pcu_sparsemat_mult_transpose_add(
        cmat->mat, alpha, ParOptCudaVec::handle(pzw), ParOptCudaVec::handle(out)));
  }
  void sync_data() {
    if (!data_dirty) return;
    const ParOptScalar *data = NULL;
    getSparseJacobianData(NULL, NULL, &data);
    PCU_ADAPTER_CHECK(pcu_sparsemat_set_data(cmat->mat, data));
    data_dirty = 0;
  }
  ParOptCudaQuasiDefSparseMat *cmat;
#endif
  int data_dirty = 1;

  void writeOutput(int iter, ParOptVec *xvec) { rec->writeOutput(iter, xvec); }

  void getVarsAndBounds(ParOptVec *xvec, ParOptVec *lbvec, ParOptVec *ubvec) {
    double *x, *lb, *ub;
    xvec->getArray(&x);
    lbvec->getArray(&lb);
    ubvec->getArray(&ub);
    uint64_t kx = stream_key(p.seed, 3);
    for (int i = 0; i < n; i++) {
      x[i] = p.x0_lo[0] + p.x0_w[0] * uniform01(kx, (uint64_t)i);
      lb[i] = p.lb[0];
      ub[i] = p.ub[0];
    }
  }
  int evalSparseObjCon(ParOptVec *xvec, ParOptScalar *fobj, ParOptScalar *cons,
                       ParOptVec *sparse_con) {
    double *x, *cw;
    xvec->getArray(&x);
    sparse_con->getArray(&cw);
    double f = 0.0;
    for (int i = 0; i < n; i++) f += 0.5 * lam[i] * x[i] * x[i] + b[i] * x[i];
    *fobj = f;
    for (int j = 0; j < p.ncon; j++) {
      uint64_t key = stream_key(p.seed, 100 + (uint64_t)j);
      double sum = 0.0;
      for (int i = 0; i < n; i++) sum += (p.a_lo + p.a_w * uniform01(key, (uint64_t)i)) * x[i];
      cons[j] = beta[j] + sum;
    }
    const int *rowp, *cols;
    getSparseJacobianData(&rowp, &cols, NULL);
    for (int i = 0; i < W; i++) {
      double sum = 0.0;
      for (int e = rowp[i]; e < rowp[i + 1]; e++) {
        const double d = x[cols[e]] - xc();
        sum += wk[e] * d * d;
      }
      cw[i] = rad() - sum;
    }
    return 0;
  }
  int evalSparseObjConGradient(ParOptVec *xvec, ParOptVec *gvec, ParOptVec **Ac,
                               ParOptScalar *data) {
    double *x, *g;
    xvec->getArray(&x);
    gvec->getArray(&g);
    for (int i = 0; i < n; i++) g[i] = lam[i] * x[i] + b[i];
    for (int j = 0; j < p.ncon; j++) {
      double *a;
      Ac[j]->getArray(&a);
      uint64_t key = stream_key(p.seed, 100 + (uint64_t)j);
      for (int i = 0; i < n; i++) a[i] = p.a_lo + p.a_w * uniform01(key, (uint64_t)i);
    }
    const int *rowp, *cols;
    getSparseJacobianData(&rowp, &cols, NULL);
    for (int i = 0; i < W; i++)
      for (int e = rowp[i]; e < rowp[i + 1]; e++) data[e] = -2.0 * wk[e] * (x[cols[e]] - xc());
    data_dirty = 1;
    return 0;
  }

  SepQuadParams p;
  Recorder *rec;
  int n, W;
  std::vector<double> wk, lam, b, beta;
};

// ---------------------------------------------------------------------------
// rosenbrock: same mathematical problem as the class in
// /root/reference/examples/rosenbrock/rosenbrock.cpp:9-199 (scale = 1)
// ---------------------------------------------------------------------------
class Rosen : public HistoryProblem {
 public:
  Rosen(MPI_Comm comm, int nvars_, int nwcon_, int nwstart_, int nw_,
        int nwskip_)
      : HistoryProblem(comm) {
    n = nvars_;
    nwc = nwcon_;
    nwstart = nwstart_;
    nw = nw_;
    nwskip = nwskip_;
    setProblemSizes(n, 2, nwc);
    setNumInequalities(2, nwc);
  }
  ParOptQuasiDefMat *createQuasiDefMat() {
#ifdef PCU_ADAPTERS
    pcu_weighting w;
    memset(&w, 0, sizeof(w));
    w.nwcon = nwc;
    w.wstart = nwstart;
    w.nw = nw;
    w.wstride = nw + nwskip;
    w.coef0 = -1.0;
    w.coef_rest = -1.0;
    w.wconst = 1.0;
    return new ParOptCudaQuasiDefBlockMat(ParOptCudaContext(), n, &w);
#else
    return new ParOptQuasiDefBlockMat(this, 1);
#endif
  }
  void getVarsAndBounds(ParOptVec *xvec, ParOptVec *lbvec, ParOptVec *ubvec) {
    xvec->set(-1.0);
    lbvec->set(-2.0);
    ubvec->set(1.0);
  }
  int evalObjCon(ParOptVec *xvec, ParOptScalar *fobj, ParOptScalar *cons) {
    double *x;
    xvec->getArray(&x);
    double obj = 0.0, c0 = 0.0, c1 = 0.0;
    for (int i = 0; i < n - 1; i++) {
      double d = x[i + 1] - x[i] * x[i];
      obj += (1.0 - x[i]) * (1.0 - x[i]) + 100.0 * d * d;
    }
    for (int i = 0; i < n; i++) c0 -= x[i] * x[i];
    for (int i = 0; i < n; i += 2) c1 += x[i];
    *fobj = obj;
    cons[0] = c0 + 0.25;
    cons[1] = c1 + 10.0;
    return 0;
  }
  int evalObjConGradient(ParOptVec *xvec, ParOptVec *gvec, ParOptVec **Ac) {
    double *x, *g, *a0, *a1;
    xvec->getArray(&x);
    gvec->getArray(&g);
    gvec->zeroEntries();
    for (int i = 0; i < n - 1; i++) {
      double d = x[i + 1] - x[i] * x[i];
      g[i] += -2.0 * (1.0 - x[i]) + 200.0 * d * (-2.0 * x[i]);
      g[i + 1] += 200.0 * d;
    }
    Ac[0]->getArray(&a0);
    Ac[1]->getArray(&a1);
    for (int i = 0; i < n; i++) a0[i] = -2.0 * x[i];
    for (int i = 0; i < n; i++) a1[i] = (i % 2 == 0) ? 1.0 : 0.0;
    return 0;
  }
  void evalSparseCon(ParOptVec *xvec, ParOptVec *out) {
    double *x, *o;
    xvec->getArray(&x);
    out->getArray(&o);
    for (int i = 0, j = nwstart; i < nwc; i++, j += nwskip) {
      o[i] = 1.0;
      for (int k = 0; k < nw; k++, j++) o[i] -= x[j];
    }
  }
  void addSparseJacobian(ParOptScalar alpha, ParOptVec *, ParOptVec *px,
                         ParOptVec *out) {
    double *v, *o;
    px->getArray(&v);
    out->getArray(&o);
    for (int i = 0, j = nwstart; i < nwc; i++, j += nwskip) {
      for (int k = 0; k < nw; k++, j++) o[i] -= alpha * v[j];
    }
  }
  void addSparseJacobianTranspose(ParOptScalar alpha, ParOptVec *,
                                  ParOptVec *pzw, ParOptVec *out) {
    double *z, *o;
    pzw->getArray(&z);
    out->getArray(&o);
    for (int i = 0, j = nwstart; i < nwc; i++, j += nwskip) {
      for (int k = 0; k < nw; k++, j++) o[j] -= alpha * z[i];
    }
  }
  void addSparseInnerProduct(ParOptScalar alpha, ParOptVec *, ParOptVec *cvec,
                             ParOptScalar *A) {
    double *cv;
    cvec->getArray(&cv);
    for (int i = 0, j = nwstart; i < nwc; i++, j += nwskip) {
      for (int k = 0; k < nw; k++, j++) A[i] += alpha * cv[j];
    }
  }
  int n, nwc, nwstart, nw, nwskip;
};

// ---------------------------------------------------------------------------
static bool arg_d(const char *a, const char *key, double *v) {
  size_t k = strlen(key);
  if (strncmp(a, key, k) == 0 && a[k] == '=') {
    *v = atof(a + k + 1);
    return true;
  }
  return false;
}

int main(int argc, char *argv[]) {
  MPI_Init(&argc, &argv);
  MPI_Comm comm = MPI_COMM_WORLD;
  int rank;
  MPI_Comm_rank(comm, &rank);

  std::string problem = "sepquad", hist_path, out_path = "/dev/null";
  std::string algorithm = "ip", tr_log = "/dev/null";
  std::string dump_x, checkpoint;
  SepQuadParams p;
  std::vector<std::pair<std::string, std::string> > opts;
  int rosen_n = 1000;
  for (int k = 1; k < argc; k++) {
    const char *a = argv[k];
    double v;
    if (strncmp(a, "problem=", 8) == 0) problem = a + 8;
    else if (strncmp(a, "hist=", 5) == 0) hist_path = a + 5;
    else if (strncmp(a, "log=", 4) == 0) out_path = a + 4;
    else if (strncmp(a, "dump_x=", 7) == 0) dump_x = a + 7;
    else if (strncmp(a, "checkpoint=", 11) == 0) checkpoint = a + 11;
    else if (strncmp(a, "algorithm=", 10) == 0) algorithm = a + 10;
    else if (strncmp(a, "tr_log=", 7) == 0) tr_log = a + 7;
    else if (arg_d(a, "n", &v)) { p.ntotal = (long)v; rosen_n = (int)v; }
    else if (arg_d(a, "ncon", &v)) p.ncon = (int)v;
    else if (arg_d(a, "nw", &v)) p.nw = (int)v;
    else if (arg_d(a, "nb", &v)) p.nb = (int)v;
    else if (arg_d(a, "seed", &v)) p.seed = (uint64_t)v;
    else if (arg_d(a, "lam_min", &v)) p.lam_min = v;
    else if (arg_d(a, "lam_max", &v)) p.lam_max = v;
    else if (arg_d(a, "b_lo", &v)) p.b_lo = v;
    else if (arg_d(a, "b_w", &v)) p.b_w = v;
    else if (arg_d(a, "a_lo", &v)) p.a_lo = v;
    else if (arg_d(a, "a_w", &v)) p.a_w = v;
    else if (arg_d(a, "beta_c", &v)) p.beta_c = v;
    else if (arg_d(a, "beta_n", &v)) p.beta_n = v;
    else if (arg_d(a, "beta_u", &v)) p.beta_u = v;
    else if (arg_d(a, "x0_lo0", &v)) p.x0_lo[0] = v;
    else if (arg_d(a, "x0_lo1", &v)) p.x0_lo[1] = v;
    else if (arg_d(a, "x0_w0", &v)) p.x0_w[0] = v;
    else if (arg_d(a, "x0_w1", &v)) p.x0_w[1] = v;
    else if (arg_d(a, "lb0", &v)) p.lb[0] = v;
    else if (arg_d(a, "lb1", &v)) p.lb[1] = v;
    else if (arg_d(a, "ub0", &v)) p.ub[0] = v;
    else if (arg_d(a, "ub1", &v)) p.ub[1] = v;
    else if (arg_d(a, "householder", &v)) p.householder = (int)v;
    else if (strncmp(a, "opt:", 4) == 0) {
      // opt:name=value  -> forwarded to ParOptOptions
      const char *eq = strchr(a + 4, '=');
      if (eq) opts.push_back({std::string(a + 4, eq - (a + 4)), eq + 1});
    }
  }

  HistoryProblem *prob = NULL;      // history hook + bookkeeping
  ParOptProblem *opt_prob = NULL;   // what the optimizer sees
  if (problem == "rosenbrock") {
    // rosenbrock.cpp:225-229: nvars-1 variables, nwcon=5, nw=5, start 1, skip 1
    prob = new Rosen(comm, rosen_n - 1, 5, 1, 5, 1);
  } else if (problem == "sparsequad") {
    SparseQuad *sq = new SparseQuad(comm, p);
    sq->incref();
    prob = sq->rec;
    opt_prob = sq;
  } else {
    prob = new SepQuad(comm, p);
  }
  prob->incref();
  if (!opt_prob) opt_prob = prob;

  ParOptOptions *options = new ParOptOptions(comm);
  if (algorithm == "tr") ParOptOptimizer::addDefaultOptions(options);
  else ParOptInteriorPoint::addDefaultOptions(options);
  options->incref();
  options->setOption("output_file", out_path.c_str());
  options->setOption("write_output_frequency", 1);
  if (algorithm == "tr") {
    options->setOption("algorithm", "tr");
    options->setOption("tr_output_file", tr_log.c_str());
    options->setOption("tr_write_output_frequency", 1);
    options->setOption("output_level", 0);
  }
  for (size_t i = 0; i < opts.size(); i++) {
    const char *name = opts[i].first.c_str();
    const char *val = opts[i].second.c_str();
    if (options->isOption(name)) {
      int type = options->getOptionType(name);
      int fail = 1;
      if (type == ParOptOptions::PAROPT_INT_OPTION ||
          type == ParOptOptions::PAROPT_BOOLEAN_OPTION) {
        fail = options->setOption(name, atoi(val));
      } else if (type == ParOptOptions::PAROPT_FLOAT_OPTION) {
        fail = options->setOption(name, atof(val));
      } else {
        fail = options->setOption(name, val);
      }
      if (fail && rank == 0) {
        fprintf(stderr, "ref_driver: failed to set option %s=%s\n", name, val);
      }
    } else if (rank == 0) {
      fprintf(stderr, "ref_driver: unknown option %s\n", name);
    }
  }

  if (algorithm == "tr") {
    // ParOptOptimizer with the trust-region front end (ParOptOptimizer.cpp:102-175,
    // the configuration of examples/rosenbrock/rosenbrock.cpp:234-242)
    prob->tr_mode = 1;
    if (rank == 0 && !hist_path.empty()) prob->hist = fopen(hist_path.c_str(), "w");
    ParOptOptimizer *opt = new ParOptOptimizer(opt_prob, options);
    opt->incref();
    double t0 = MPI_Wtime();
    opt->optimize();
    double t1 = MPI_Wtime();
    ParOptVec *x;
    opt->getOptimizedPoint(&x, NULL, NULL, NULL, NULL);
    double xs = vsum(x), xn = x->norm();
    if (rank == 0) {
      FILE *fp = prob->hist ? prob->hist : stdout;
      fprintf(fp, "{\"final\": 1, \"xsum\": %.17g, \"xnorm\": %.17g, \"time_s\": %.6f}\n", xs,
              xn, t1 - t0);
      if (prob->hist) fclose(prob->hist);
    }
    opt->decref();
    options->decref();
    prob->decref();
    MPI_Finalize();
    return 0;
  }

  ParOptInteriorPoint *ip = new ParOptInteriorPoint(opt_prob, options);
  ip->incref();
  prob->ip = ip;
#ifdef PCU_ADAPTERS
  {
    // the compact quasi-Newton object of the options, on the device
    const char *qn_type = options->getEnumOption("qn_type");
    if (strcmp(qn_type, "bfgs") == 0 || strcmp(qn_type, "sr1") == 0) {
      int nv, nc, nw;
      opt_prob->getProblemSizes(&nv, &nc, &nw);
      ParOptCudaCompactQN *cqn = new ParOptCudaCompactQN(
          ParOptCudaContext(), nv, qn_type, options->getIntOption("qn_subspace_size"));
      if (strcmp(options->getEnumOption("qn_update_type"), "damped_update") == 0)
        cqn->setBFGSUpdateType(PAROPT_DAMPED_UPDATE);
      if (strcmp(options->getEnumOption("qn_diag_type"), "yts_over_sts") == 0)
        cqn->setInitDiagonalType(PAROPT_YTS_OVER_STS);
      ip->setQuasiNewton(cqn);
    }
    if (rank == 0) {
      double *probe;
      ip->variables.x->getArray(&probe);
      fprintf(stderr, "adapter_driver: reference ParOptInteriorPoint on %s vectors\n",
              dynamic_cast<ParOptCudaVec *>(ip->variables.x) ? "ParOptCudaVec" : "HOST");
    }
  }
#endif
  if (rank == 0 && !hist_path.empty()) prob->hist = fopen(hist_path.c_str(), "w");

  double t0 = MPI_Wtime();
  int fail = ip->optimize();
  double t1 = MPI_Wtime();

  // the reference's own binary checkpoint of the final state (IP.cpp:883-975)
  if (!checkpoint.empty()) ip->writeSolutionFile(checkpoint.c_str());

  int niter, neval, ngeval;
  ip->getIterationCounters(&niter, &neval, &ngeval);
  if (rank == 0) {
    FILE *fp = prob->hist ? prob->hist : stdout;
    fprintf(fp,
            "{\"final\": 1, \"fail\": %d, \"niter\": %d, \"neval\": %d, "
            "\"ngeval\": %d, \"fobj\": %.17g, \"mu\": %.17g, \"time_s\": %.6f, "
            "\"callback_s\": %.6f}\n",
            fail, niter, neval, ngeval, ip->fobj, ip->barrier_param, t1 - t0,
            prob->callback_time);
    if (prob->hist) fclose(prob->hist);
    if (!dump_x.empty()) {
      ParOptVec *x;
      ip->getOptimizedPoint(&x, NULL, NULL, NULL, NULL);
      double *xa;
      int n = x->getArray(&xa);
      FILE *fx = fopen(dump_x.c_str(), "wb");
      if (fx) {
        fwrite(xa, sizeof(double), n, fx);
        fclose(fx);
      }
    }
  }

#ifdef PCU_ADAPTERS
  if (rank == 0)
    fprintf(stderr, "adapter_driver: %lld kernels of libparopt_b200 launched\n",
            (long long)pcu_ctx_kernel_launches(ParOptCudaContext()));
#endif
  ip->decref();
  options->decref();
  prob->decref();
  MPI_Finalize();
  return 0;
}
