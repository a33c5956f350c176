// pcu_host_sepquad.cu -- the separable / Householder QP workload (DESIGN.md
// "Synthetic problems") written the way a user of the reference writes a C++
// ParOptProblem: callbacks over host arrays (ParOptVec::getArray), threaded over
// the host cores.  It is a CLIENT of the public C ABI (pcu_problem_create_host,
// pcu_ctx_allreduce_sum): the end-to-end benchmark runs the optimizer through
// it, so that the iterate crosses PCIe device->host and the gradients
// host->device on every callback.  Counterpart of the problem class in
// oracle/ref_driver.cpp (same generator, same arithmetic).
#include <emmintrin.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../include/paropt_b200.h"

namespace {

inline uint64_t hs_splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
inline uint64_t hs_key(uint64_t seed, uint64_t stream) {
  return hs_splitmix64(seed ^ (stream * 0x9E3779B97F4A7C15ULL));
}
inline double hs_u01(uint64_t key, uint64_t idx) {
  return (double)(hs_splitmix64(key + idx) >> 11) * (1.0 / 9007199254740992.0);
}

struct HostSepQuad {
  pcu_ctx *ctx = nullptr;
  pcu_sepquad_params p;
  long long offset = 0;
  int n = 0, ncon = 0, nthreads = 1;
  std::vector<double> lam, b, vh;
  std::vector<std::vector<double> > a;  // constraint gradients (constant)
  std::vector<double> beta;
  double vtv = 0.0;

  // fixed static partition: thread t owns [n t / T, n (t+1) / T) -- the sums are
  // combined in thread order, so results do not depend on scheduling
  template <class Body>
  void parallel(const Body &body) const {
    const int T = nthreads;
    std::vector<std::thread> th;
    th.reserve(T > 0 ? T - 1 : 0);
    for (int t = 1; t < T; t++) {
      th.emplace_back([&, t]() {
        body(t, (long long)n * t / T, (long long)n * (t + 1) / T);
      });
    }
    body(0, 0, (long long)n / T);
    for (auto &x : th) x.join();
  }

  int allreduce(double *v, int k) const { return pcu_ctx_allreduce_sum(ctx, v, k); }

  void init() {
    lam.resize(n);
    b.resize(n);
    if (p.householder) vh.resize(n);
    a.assign(ncon, std::vector<double>());
    for (int j = 0; j < ncon; j++) a[j].resize(n);
    const uint64_t klam = hs_key(p.seed, 1), kb = hs_key(p.seed, 2), kv = hs_key(p.seed, 7);
    std::vector<uint64_t> ka(ncon);
    for (int j = 0; j < ncon; j++) ka[j] = hs_key(p.seed, 100 + (uint64_t)j);
    std::vector<double> part(nthreads, 0.0);
    parallel([&](int t, long long lo, long long hi) {
      double s = 0.0;
      for (long long i = lo; i < hi; i++) {
        const uint64_t gi = (uint64_t)(offset + i);
        lam[i] = p.lam_min + (p.lam_max - p.lam_min) * hs_u01(klam, gi);
        b[i] = p.b_lo + p.b_w * hs_u01(kb, gi);
        if (p.householder) {
          vh[i] = 0.5 + hs_u01(kv, gi);
          s += vh[i] * vh[i];
        }
        for (int j = 0; j < ncon; j++) a[j][i] = p.a_lo + p.a_w * hs_u01(ka[j], gi);
      }
      part[t] = s;
    });
    vtv = 0.0;
    for (double s : part) vtv += s;
    allreduce(&vtv, 1);
    beta.resize(ncon);
    const uint64_t kbeta = hs_key(p.seed, 5);
    for (int j = 0; j < ncon; j++)
      beta[j] = p.beta_c + p.beta_n * (double)p.ntotal + p.beta_u * hs_u01(kbeta, (uint64_t)j);
  }

  double householder_factor(const double *x) const {  // 2 (v.x) / (v.v)
    if (!p.householder) return 0.0;
    std::vector<double> part(nthreads, 0.0);
    parallel([&](int t, long long lo, long long hi) {
      double s = 0.0;
      for (long long i = lo; i < hi; i++) s += vh[i] * x[i];
      part[t] = s;
    });
    double vx = 0.0;
    for (double s : part) vx += s;
    allreduce(&vx, 1);
    return 2.0 * vx / vtv;
  }

  static int get_vars(void *user, int n, double *x, double *lb, double *ub) {
    HostSepQuad *q = (HostSepQuad *)user;
    const uint64_t kx = hs_key(q->p.seed, 3);
    q->parallel([&](int, long long lo, long long hi) {
      for (long long i = lo; i < hi; i++) {
        const int k = (q->p.nw > 0 && (i % q->p.nw) != 0) ? 1 : 0;
        x[i] = q->p.x0_lo[k] + q->p.x0_w[k] * hs_u01(kx, (uint64_t)(q->offset + i));
        lb[i] = q->p.lb[k];
        ub[i] = q->p.ub[k];
      }
    });
    (void)n;
    return 0;
  }

  static int eval_obj(void *user, int n, const double *x, double *fobj, double *cons) {
    HostSepQuad *q = (HostSepQuad *)user;
    const int c = q->ncon, T = q->nthreads;
    const double hf = q->householder_factor(x);
    std::vector<double> part((size_t)T * (c + 1), 0.0);
    q->parallel([&](int t, long long lo, long long hi) {
      // one pass over x: the objective and the c constraint products together
      const double *lam = q->lam.data(), *b = q->b.data();
      const double *vh = q->p.householder ? q->vh.data() : nullptr;
      constexpr int CB = 16;  // constraints per sweep (c <= 16: a single sweep)
      double f = 0.0;
      for (int j0 = 0; j0 < c || j0 == 0; j0 += CB) {
        const int cb = c - j0 < CB ? c - j0 : CB;
        const double *aj[CB];
        double sj[CB];
        for (int j = 0; j < cb; j++) {
          aj[j] = q->a[j0 + j].data();
          sj[j] = 0.0;
        }
        for (long long i = lo; i < hi; i++) {
          const double xi = x[i];
          if (j0 == 0) {
            if (vh) {
              const double y = xi - hf * vh[i];
              f += 0.5 * lam[i] * y * y + b[i] * xi;
            } else {
              f += (0.5 * lam[i] * xi + b[i]) * xi;
            }
          }
          for (int j = 0; j < cb; j++) sj[j] += aj[j][i] * xi;
        }
        for (int j = 0; j < cb; j++) part[(size_t)t * (c + 1) + 1 + j0 + j] = sj[j];
        if (c == 0) break;
      }
      part[(size_t)t * (c + 1)] = f;
    });
    std::vector<double> tot(c + 1, 0.0);
    for (int t = 0; t < T; t++)
      for (int j = 0; j <= c; j++) tot[j] += part[(size_t)t * (c + 1) + j];
    if (q->allreduce(tot.data(), c + 1)) return 1;
    *fobj = tot[0];
    for (int j = 0; j < c; j++) cons[j] = q->beta[j] + tot[1 + j];
    (void)n;
    return 0;
  }

  static int eval_grad(void *user, int n, const double *x, double *g, double **Ac) {
    HostSepQuad *q = (HostSepQuad *)user;
    const int c = q->ncon, T = q->nthreads;
    const double *lam = q->lam.data(), *b = q->b.data();
    if (q->p.householder) {
      const double *vh = q->vh.data();
      const double hf = q->householder_factor(x);
      std::vector<double> part(T, 0.0);
      q->parallel([&](int t, long long lo, long long hi) {
        double s = 0.0;
        for (long long i = lo; i < hi; i++) {
          g[i] = lam[i] * (x[i] - hf * vh[i]);
          s += vh[i] * g[i];
        }
        part[t] = s;
      });
      double vw = 0.0;
      for (double s : part) vw += s;
      if (q->allreduce(&vw, 1)) return 1;
      const double hf2 = 2.0 * vw / q->vtv;
      q->parallel([&](int, long long lo, long long hi) {
        for (long long i = lo; i < hi; i++) g[i] = (g[i] - hf2 * vh[i]) + b[i];
      });
    } else {
      q->parallel([&](int, long long lo, long long hi) {
        // g is write-only here: streaming stores skip the read-for-ownership
        long long i = lo;
        for (; i < hi && (((uintptr_t)(g + i)) & 15); i++) g[i] = lam[i] * x[i] + b[i];
        for (; i + 2 <= hi; i += 2)
          _mm_stream_pd(g + i, _mm_set_pd(lam[i + 1] * x[i + 1] + b[i + 1], lam[i] * x[i] + b[i]));
        for (; i < hi; i++) g[i] = lam[i] * x[i] + b[i];
        _mm_sfence();
      });
    }
    // the constraints are linear: the gradients are the stored coefficient rows
    q->parallel([&](int, long long lo, long long hi) {
      for (int j = 0; j < c; j++)
        memcpy(Ac[j] + lo, q->a[j].data() + lo, sizeof(double) * (size_t)(hi - lo));
    });
    (void)n;
    return 0;
  }
};

}  // namespace

extern "C" {

pcu_problem *pcu_problem_create_sepquad_host(pcu_ctx *ctx,
                                             const pcu_sepquad_params *params,
                                             int nthreads, void **user_out) {
  if (!ctx || !params) return nullptr;
  HostSepQuad *q = new HostSepQuad;
  q->ctx = ctx;
  q->p = *params;
  const int rank = pcu_ctx_rank(ctx), world = pcu_ctx_size(ctx);
  const long long unit = params->nw > 0 ? params->nw : 1;
  const long long nunits = params->ntotal / unit;
  const long long u0 = (nunits * rank) / world, u1 = (nunits * (rank + 1)) / world;
  q->offset = u0 * unit;
  long long n = (u1 - u0) * unit;
  if (rank == world - 1) n = params->ntotal - q->offset;
  q->n = (int)n;
  q->ncon = params->ncon;
  if (nthreads < 1) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  q->nthreads = nthreads;
  q->init();
  pcu_weighting w;
  memset(&w, 0, sizeof(w));
  if (params->nw > 0) {
    w.nwcon = (int)(u1 - u0);
    w.wstart = 0;
    w.nw = params->nw;
    w.wstride = params->nw;
    w.coef0 = 1.0;
    w.coef_rest = -1.0;
    w.wconst = 0.0;
  }
  pcu_host_callbacks cb;
  memset(&cb, 0, sizeof(cb));
  cb.user = q;
  cb.get_vars_and_bounds = HostSepQuad::get_vars;
  cb.eval_obj_con = HostSepQuad::eval_obj;
  cb.eval_obj_con_gradient = HostSepQuad::eval_grad;
  pcu_problem *prob = pcu_problem_create_host(ctx, q->n, q->ncon, -1, -1, 1, 1, &w, &cb);
  if (!prob) {
    delete q;
    return nullptr;
  }
  if (user_out) *user_out = q;
  return prob;
}

void pcu_problem_sepquad_host_free(void *user) { delete (HostSepQuad *)user; }

}  // extern "C"
