/*
  tests/adapters/paropt_cuda_adapters.h -- the reference-side binding of the C ABI,
  COMPILED (INTEGRATION.md shows the same classes as the patch a ParOpt maintainer
  adds next to src/ParOptVec.h).

  Adapter classes over include/paropt_b200.h for the three abstract classes of the
  reference's drop-in boundary:

    ParOptCudaVec               : ParOptVec                 (src/ParOptVec.h:53-70)
    ParOptCudaQuasiDefBlockMat  : ParOptQuasiDefMat         (src/ParOptSparseMat.h:18-62)
    ParOptCudaQuasiDefSparseMat : ParOptQuasiDefMat         (src/ParOptSparseMat.h:106-188, CSR)
    ParOptCudaCompactQN         : ParOptCompactQuasiNewton  (src/ParOptQuasiNewton.h:32-67)

  and the three factory overrides a problem class adds (src/ParOptProblem.h:58,65,72).
  With them the UNMODIFIED reference ParOptInteriorPoint runs with every BLAS-1
  operation, every reduction, the block-diagonal Ew factor / solve and the compact
  quasi-Newton algebra on the GPU.  Its ~25 private methods that loop over raw
  getArray() pointers keep working because the vectors live in unified memory
  (context parameter "managed_vectors"): pcu_vec_host_ptr synchronises the stream
  and hands out the pointer; the next library call on that vector moves it back, so a
  write through the pointer is never lost.

  This header includes reference headers: it is built only where /root/reference
  exists (oracle/Makefile target _ref/adapter_driver); the binary travels to the GPU
  box.  Test infrastructure: nothing under paropt_b200/ includes it.
*/
#ifndef PAROPT_CUDA_ADAPTERS_H
#define PAROPT_CUDA_ADAPTERS_H

#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "ParOptProblem.h"
#include "ParOptQuasiNewton.h"
#include "ParOptSparseMat.h"
#include "ParOptVec.h"
#include "paropt_b200.h"

#define PCU_ADAPTER_CHECK(call)                                             \
  do {                                                                      \
    if ((call) != 0) {                                                      \
      fprintf(stderr, "paropt_cuda_adapters: %s failed (%s:%d)\n", #call,   \
              __FILE__, __LINE__);                                          \
      abort();                                                              \
    }                                                                       \
  } while (0)

/* One context per rank (replaces the MPI_Comm of ParOptBasicVec). */
inline pcu_ctx *ParOptCudaContext() {
  static pcu_ctx *ctx = NULL;
  if (!ctx) {
    ctx = pcu_ctx_create(0);
    if (!ctx) {
      fprintf(stderr, "paropt_cuda_adapters: no CUDA device\n");
      abort();
    }
    PCU_ADAPTER_CHECK(pcu_ctx_set_param(ctx, "managed_vectors", 1));
  }
  return ctx;
}

class ParOptCudaVec : public ParOptVec {
 public:
  ParOptCudaVec(pcu_ctx *ctx, int n) : owns(1) {
    v = pcu_vec_create(ctx, n);
    if (!v) abort();
  }
  /* borrowed handle (the Z vectors of a pcu_qn) */
  explicit ParOptCudaVec(pcu_vec *handle) : v(handle), owns(0) {}
  ~ParOptCudaVec() {
    if (owns) pcu_vec_destroy(v);
  }
  static pcu_vec *handle(ParOptVec *vec) {
    ParOptCudaVec *c = dynamic_cast<ParOptCudaVec *>(vec);
    return c ? c->v : NULL;
  }

  void set(ParOptScalar alpha) { PCU_ADAPTER_CHECK(pcu_vec_set(v, alpha)); }
  void zeroEntries() { PCU_ADAPTER_CHECK(pcu_vec_zero(v)); }
  void copyValues(ParOptVec *vec) {
    pcu_vec *o = handle(vec);
    if (o) PCU_ADAPTER_CHECK(pcu_vec_copy(v, o)); /* foreign type: ignored, ParOptVec.cpp:51-55 */
  }
  double norm() {
    double r = 0.0;
    PCU_ADAPTER_CHECK(pcu_vec_norm(v, &r));
    return r;
  }
  double maxabs() {
    double r = 0.0;
    PCU_ADAPTER_CHECK(pcu_vec_maxabs(v, &r));
    return r;
  }
  double l1norm() {
    double r = 0.0;
    PCU_ADAPTER_CHECK(pcu_vec_l1norm(v, &r));
    return r;
  }
  ParOptScalar dot(ParOptVec *vec) {
    double r = 0.0;
    pcu_vec *o = handle(vec);
    if (o) PCU_ADAPTER_CHECK(pcu_vec_dot(v, o, &r));
    return r;
  }
  void mdot(ParOptVec **vecs, int nvecs, ParOptScalar *output) {
    std::vector<pcu_vec *> hs(nvecs > 0 ? nvecs : 1);
    for (int i = 0; i < nvecs; i++) {
      hs[i] = handle(vecs[i]);
      if (!hs[i]) abort();
    }
    if (nvecs > 0) PCU_ADAPTER_CHECK(pcu_vec_mdot(v, hs.data(), nvecs, output));
  }
  void scale(ParOptScalar alpha) { PCU_ADAPTER_CHECK(pcu_vec_scale(v, alpha)); }
  void axpy(ParOptScalar alpha, ParOptVec *x) {
    pcu_vec *o = handle(x);
    if (o) PCU_ADAPTER_CHECK(pcu_vec_axpy(v, alpha, o));
  }
  int getArray(ParOptScalar **array) {
    if (array) {
      *array = pcu_vec_host_ptr(v);
      if (!*array && pcu_vec_size(v) > 0) abort();
    }
    return pcu_vec_size(v);
  }

  pcu_vec *v;

 private:
  int owns;
};

/* ParOptQuasiDefBlockMat for sparse constraints declared as a pcu_weighting. */
class ParOptCudaQuasiDefBlockMat : public ParOptQuasiDefMat {
 public:
  ParOptCudaQuasiDefBlockMat(pcu_ctx *ctx, int nvars, const pcu_weighting *w) {
    mat = pcu_blockmat_create(ctx, nvars, w);
    if (!mat) abort();
  }
  /* nwblock > 1: groups of nb rows with a dense coefficient matrix */
  ParOptCudaQuasiDefBlockMat(pcu_ctx *ctx, int nvars, const pcu_block_weighting *b) {
    mat = pcu_blockmat_create_blocks(ctx, nvars, b);
    if (!mat) abort();
  }
  ~ParOptCudaQuasiDefBlockMat() { pcu_blockmat_destroy(mat); }
  int factor(ParOptVec *x, ParOptVec *Dinv, ParOptVec *Cdiag) {
    return pcu_blockmat_factor(mat, ParOptCudaVec::handle(x), ParOptCudaVec::handle(Dinv),
                               ParOptCudaVec::handle(Cdiag));
  }
  void apply(ParOptVec *bx, ParOptVec *yx, ParOptVec *yw) {
    PCU_ADAPTER_CHECK(pcu_blockmat_apply3(mat, ParOptCudaVec::handle(bx),
                                          ParOptCudaVec::handle(yx), ParOptCudaVec::handle(yw)));
  }
  void apply(ParOptVec *bx, ParOptVec *bw, ParOptVec *yx, ParOptVec *yw) {
    PCU_ADAPTER_CHECK(pcu_blockmat_apply4(mat, ParOptCudaVec::handle(bx),
                                          ParOptCudaVec::handle(bw), ParOptCudaVec::handle(yx),
                                          ParOptCudaVec::handle(yw)));
  }
  const char *getFactorInfo() { return "paropt_b200 block-diagonal Ew (CUDA)"; }

 private:
  pcu_blockmat *mat;
};

/* ParOptQuasiDefSparseMat (src/ParOptSparseMat.cpp:231-451) on the device, for a
   ParOptSparseProblem (general CSR sparse constraints): what its createQuasiDefMat()
   override returns.  Takes the Jacobian values the problem holds at every factor(), like
   the reference (ParOptSparseMat.cpp:318-321). */
class ParOptCudaQuasiDefSparseMat : public ParOptQuasiDefMat {
 public:
  ParOptCudaQuasiDefSparseMat(pcu_ctx *ctx, ParOptSparseProblem *problem) : prob(problem) {
    int nv, nc, nw;
    prob->getProblemSizes(&nv, &nc, &nw);
    const int *rowp = NULL, *cols = NULL;
    prob->getSparseJacobianData(&rowp, &cols, NULL);
    mat = pcu_sparsemat_create(ctx, nv, nw, rowp, cols, 1);
    if (!mat) abort();
  }
  ~ParOptCudaQuasiDefSparseMat() { pcu_sparsemat_destroy(mat); }
  int factor(ParOptVec *x, ParOptVec *Dinv, ParOptVec *Cdiag) {
    const ParOptScalar *data = NULL;
    prob->getSparseJacobianData(NULL, NULL, &data);
    PCU_ADAPTER_CHECK(pcu_sparsemat_set_data(mat, data));
    return pcu_sparsemat_factor(mat, ParOptCudaVec::handle(x), ParOptCudaVec::handle(Dinv),
                                ParOptCudaVec::handle(Cdiag));
  }
  void apply(ParOptVec *bx, ParOptVec *yx, ParOptVec *yw) {
    PCU_ADAPTER_CHECK(pcu_sparsemat_apply3(mat, ParOptCudaVec::handle(bx),
                                           ParOptCudaVec::handle(yx), ParOptCudaVec::handle(yw)));
  }
  void apply(ParOptVec *bx, ParOptVec *bw, ParOptVec *yx, ParOptVec *yw) {
    PCU_ADAPTER_CHECK(pcu_sparsemat_apply4(mat, ParOptCudaVec::handle(bx),
                                           ParOptCudaVec::handle(bw), ParOptCudaVec::handle(yx),
                                           ParOptCudaVec::handle(yw)));
  }
  const char *getFactorInfo() {
    int nk = 0, nl = 0, nlev = 0, nlaunch = 0;
    pcu_sparsemat_info(mat, &nk, &nl, &nlev, &nlaunch);
    snprintf(info, sizeof(info), "paropt_b200 sparse Cholesky (CUDA) nnz(K) %d nnz(L) %d levels %d",
             nk, nl, nlev);
    return info;
  }
  pcu_sparsemat *mat;

 private:
  ParOptSparseProblem *prob;
  char info[160];
};
/* ParOptLBFGS / ParOptLSR1 on the device ("bfgs" | "sr1"). */
class ParOptCudaCompactQN : public ParOptCompactQuasiNewton {
 public:
  ParOptCudaCompactQN(pcu_ctx *ctx, int nvars, const char *qn_type, int subspace) {
    qn = pcu_qn_create(ctx, nvars, qn_type, subspace);
    if (!qn) abort();
    const int q = 2 * subspace + 1;
    d0.resize(q);
    M.resize((size_t)q * q);
    zh.resize(q);
    zv.resize(q, NULL);
  }
  ~ParOptCudaCompactQN() {
    for (size_t i = 0; i < zv.size(); i++)
      if (zv[i]) zv[i]->decref();
    pcu_qn_destroy(qn);
  }
  void setInitDiagonalType(ParOptQuasiNewtonDiagonalType t) {
    PCU_ADAPTER_CHECK(pcu_qn_set_option(
        qn, "qn_diag_type", t == PAROPT_YTS_OVER_STS ? "yts_over_sts" : "yty_over_yts"));
  }
  void setBFGSUpdateType(ParOptBFGSUpdateType t) {
    PCU_ADAPTER_CHECK(pcu_qn_set_option(
        qn, "qn_update_type",
        t == PAROPT_DAMPED_UPDATE ? "damped_update" : "skip_negative_curvature"));
  }
  void reset() { PCU_ADAPTER_CHECK(pcu_qn_reset(qn)); }
  int update(ParOptVec *, const ParOptScalar *, ParOptVec *, ParOptVec *s, ParOptVec *y) {
    int type = 0;
    PCU_ADAPTER_CHECK(
        pcu_qn_update(qn, ParOptCudaVec::handle(s), ParOptCudaVec::handle(y), &type));
    return type;
  }
  void mult(ParOptVec *x, ParOptVec *y) {
    PCU_ADAPTER_CHECK(pcu_qn_mult(qn, ParOptCudaVec::handle(x), ParOptCudaVec::handle(y)));
  }
  void multAdd(ParOptScalar alpha, ParOptVec *x, ParOptVec *y) {
    PCU_ADAPTER_CHECK(
        pcu_qn_mult_add(qn, alpha, ParOptCudaVec::handle(x), ParOptCudaVec::handle(y)));
  }
  int getCompactMat(ParOptScalar *b0, const ParOptScalar **d, const ParOptScalar **M_,
                    ParOptVec ***Z) {
    double b = 0.0;
    const int q = pcu_qn_compact(qn, &b, d0.data(), M.data(), zh.data());
    if (b0) *b0 = b;
    if (d) *d = d0.data();
    if (M_) *M_ = M.data();
    if (Z) {
      for (int i = 0; i < q; i++) {
        if (!zv[i] || ParOptCudaVec::handle(zv[i]) != zh[i]) {
          if (zv[i]) zv[i]->decref();
          zv[i] = new ParOptCudaVec(zh[i]);
          zv[i]->incref();
        }
      }
      *Z = zv.data();
    }
    return q;
  }
  int getMaxLimitedMemorySize() { return pcu_qn_max_size(qn); }

  pcu_qn *qn;

 private:
  std::vector<double> d0, M;
  std::vector<pcu_vec *> zh;
  std::vector<ParOptVec *> zv;
};

#endif /* PAROPT_CUDA_ADAPTERS_H */
