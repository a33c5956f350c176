// pcu_sparse.cu -- general sparse constraints on the device:
//   pcu_sparsemat  <->  ParOptQuasiDefSparseMat (ParOptSparseMat.cpp:231-451) for the
//                       CSR Jacobian of a ParOptSparseProblem (ParOptProblem.h:301-407),
//                       plus the two CSR products of that class (ParOptProblem.cpp:756-816).
//
//   factor(x, Dinv, C):  K = C + A D^-1 A^T (sparse, symmetric positive definite), K = L L^T
//   apply(bx[, bw]):     yw = K^-1 (bw - A D^-1 bx),  yx = D^-1 (bx + A^T yw)
//
// The reference assembles K with ParOptMatMatTransNumeric and factors it with its own
// supernodal sparse Cholesky under a METIS / AMD ordering -- serial, on the host.  Here:
//   * symbolic phase once, on the host, at creation: CSC transpose, pattern of K,
//     minimum-degree ordering, elimination tree, pattern of L, row lists, level sets of
//     the elimination tree;
//   * numeric phase on the device: one thread per entry of K for the assembly (sorted
//     merge of two CSR rows); left-looking column Cholesky, one thread block per column,
//     one launch per level of the elimination tree (runs of single-column levels share
//     one launch); level-scheduled triangular solves, one warp per row / column.
// Every sum runs in a fixed order: results are reproducible run to run.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <set>
#include <utility>
#include <vector>

#include "pcu_ctx.cuh"

struct SparseLaunch {  // levels [l0, l1): one launch; serial = every level has one column
  int l0, l1, grid, serial;
};

struct pcu_sparsemat {
  pcu_ctx *ctx = nullptr;
  int nvars = 0, nwcon = 0, nnz = 0;
  // ---- host symbolic data
  std::vector<int> rowp, cols;          // CSR as given
  std::vector<int> srt_cols, srt_idx;   // per row: columns ascending + index into `data`
  std::vector<int> colp, rows, tmap;    // CSC (transpose) + index into `data`
  std::vector<int> perm, iperm;         // perm[new] = old
  std::vector<int> Lp, Li;              // L, CSC, rows ascending, first entry = diagonal
  std::vector<int> Rp, Rk, Rpos;        // strictly lower rows of L: column k, position in Li/Lx
  std::vector<int> kpos, ka, kb;        // K entries: position in Lx, the two rows of A (old numbering)
  std::vector<int> level_ptr, level_cols;
  std::vector<SparseLaunch> launches;
  int nnzK = 0;
  // ---- device data
  int *d_rowp = nullptr, *d_cols = nullptr, *d_srt_cols = nullptr, *d_srt_idx = nullptr;
  int *d_colp = nullptr, *d_rows = nullptr, *d_tmap = nullptr, *d_perm = nullptr;
  int *d_Lp = nullptr, *d_Li = nullptr, *d_Rp = nullptr, *d_Rk = nullptr, *d_Rpos = nullptr;
  int *d_kpos = nullptr, *d_ka = nullptr, *d_kb = nullptr;
  int *d_level_ptr = nullptr, *d_level_cols = nullptr, *d_fail = nullptr;
  int *d_where = nullptr;  // row -> position maps of the columns being factored (nslots x nwcon)
  int nslots = 1;
  double *d_data = nullptr, *d_Lx = nullptr, *d_work = nullptr;
  pcu_vec *Dinv = nullptr;  // kept from factor() like the reference (SM.cpp:306-312)
  int factored = 0;
};

// ----------------------------------------------------------------- symbolic
namespace {

// minimum-degree ordering on the elimination graph (explicit fill; ties by index)
void minimum_degree(int n, const std::vector<std::vector<int> > &adj, std::vector<int> &perm) {
  std::vector<std::set<int> > g(n);
  for (int i = 0; i < n; i++)
    for (int j : adj[i])
      if (j != i) g[i].insert(j);
  std::set<std::pair<int, int> > queue;
  for (int i = 0; i < n; i++) queue.insert({(int)g[i].size(), i});
  perm.clear();
  perm.reserve(n);
  std::vector<int> nb;
  while (!queue.empty()) {
    const int v = queue.begin()->second;
    queue.erase(queue.begin());
    perm.push_back(v);
    nb.assign(g[v].begin(), g[v].end());
    for (int u : nb) {
      queue.erase({(int)g[u].size(), u});
      g[u].erase(v);
    }
    for (size_t a = 0; a < nb.size(); a++)
      for (size_t b = a + 1; b < nb.size(); b++) {
        g[nb[a]].insert(nb[b]);
        g[nb[b]].insert(nb[a]);
      }
    for (int u : nb) queue.insert({(int)g[u].size(), u});
    g[v].clear();
  }
}

int symbolic(pcu_sparsemat *m, int ordering) {
  const int nw = m->nwcon, nv = m->nvars;
  const std::vector<int> &rowp = m->rowp, &cols = m->cols;
  m->nnz = rowp[nw];
  for (int e = 0; e < m->nnz; e++)
    if (cols[e] < 0 || cols[e] >= nv) return 1;
  // rows sorted by column
  m->srt_cols.resize(m->nnz);
  m->srt_idx.resize(m->nnz);
  for (int i = 0; i < nw; i++) {
    std::vector<std::pair<int, int> > r;
    for (int e = rowp[i]; e < rowp[i + 1]; e++) r.push_back({cols[e], e});
    std::sort(r.begin(), r.end());
    for (size_t t = 0; t < r.size(); t++) {
      if (t > 0 && r[t].first == r[t - 1].first) return 1;  // duplicate entry
      m->srt_cols[rowp[i] + t] = r[t].first;
      m->srt_idx[rowp[i] + t] = r[t].second;
    }
  }
  // transpose (ParOptSparseTranspose, SM.cpp:246)
  m->colp.assign(nv + 1, 0);
  for (int e = 0; e < m->nnz; e++) m->colp[cols[e] + 1]++;
  for (int k = 0; k < nv; k++) m->colp[k + 1] += m->colp[k];
  m->rows.resize(m->nnz);
  m->tmap.resize(m->nnz);
  {
    std::vector<int> next(m->colp.begin(), m->colp.end() - 1);
    for (int i = 0; i < nw; i++)
      for (int e = rowp[i]; e < rowp[i + 1]; e++) {
        const int at = next[cols[e]]++;
        m->rows[at] = i;
        m->tmap[at] = e;
      }
  }
  // adjacency of K = A A^T (+ diagonal), old numbering
  std::vector<std::vector<int> > adj(nw);
  {
    std::vector<int> mark(nw, -1);
    for (int i = 0; i < nw; i++) {
      mark[i] = i;
      adj[i].push_back(i);
      for (int e = rowp[i]; e < rowp[i + 1]; e++) {
        const int k = cols[e];
        for (int f = m->colp[k]; f < m->colp[k + 1]; f++) {
          const int j = m->rows[f];
          if (mark[j] != i) {
            mark[j] = i;
            adj[i].push_back(j);
          }
        }
      }
    }
  }
  if (ordering == 1) {
    minimum_degree(nw, adj, m->perm);
  } else {
    m->perm.resize(nw);
    for (int i = 0; i < nw; i++) m->perm[i] = i;
  }
  m->iperm.assign(nw, 0);
  for (int i = 0; i < nw; i++) m->iperm[m->perm[i]] = i;
  // lower pattern of the permuted K by columns, then the pattern of L:
  // pattern(j) = K(j) U union over children c of pattern(c) \ {c}; parent(j) = first row > j
  std::vector<std::vector<int> > Kcol(nw), Lcol(nw), children(nw);
  m->nnzK = 0;
  for (int jn = 0; jn < nw; jn++) {
    for (int i : adj[m->perm[jn]]) {
      const int in = m->iperm[i];
      if (in >= jn) Kcol[jn].push_back(in);
    }
    std::sort(Kcol[jn].begin(), Kcol[jn].end());
    m->nnzK += (int)Kcol[jn].size();
  }
  bool fill_overflow = false;
  {
    // the factor is addressed with 32-bit indices: give up beyond PCU_SPARSE_MAX_FILL
    // entries (a bad ordering of a matrix with long-range couplings fills in completely)
    const long long max_fill = 400000000LL;
    long long total = 0;
    std::vector<int> mark(nw, -1);
    for (int j = 0; j < nw && !fill_overflow; j++) {
      std::vector<int> &p = Lcol[j];
      for (int i : Kcol[j]) {
        mark[i] = j;
        p.push_back(i);
      }
      for (int c : children[j]) {
        for (int i : Lcol[c])
          if (i != c && mark[i] != j) {
            mark[i] = j;
            p.push_back(i);
          }
      }
      std::sort(p.begin(), p.end());
      if (p.size() > 1) children[p[1]].push_back(j);
      total += (long long)p.size();
      if (total > max_fill) fill_overflow = true;
    }
  }
  if (fill_overflow) return 2;
  m->Lp.assign(nw + 1, 0);
  for (int j = 0; j < nw; j++) m->Lp[j + 1] = m->Lp[j] + (int)Lcol[j].size();
  m->Li.resize(m->Lp[nw]);
  for (int j = 0; j < nw; j++) std::copy(Lcol[j].begin(), Lcol[j].end(), m->Li.begin() + m->Lp[j]);
  // row lists (columns ascending)
  m->Rp.assign(nw + 1, 0);
  for (int k = 0; k < nw; k++)
    for (int t = m->Lp[k] + 1; t < m->Lp[k + 1]; t++) m->Rp[m->Li[t] + 1]++;
  for (int j = 0; j < nw; j++) m->Rp[j + 1] += m->Rp[j];
  m->Rk.resize(m->Rp[nw]);
  m->Rpos.resize(m->Rp[nw]);
  {
    std::vector<int> next(m->Rp.begin(), m->Rp.end() - 1);
    for (int k = 0; k < nw; k++)
      for (int t = m->Lp[k] + 1; t < m->Lp[k + 1]; t++) {
        const int at = next[m->Li[t]]++;
        m->Rk[at] = k;
        m->Rpos[at] = t;
      }
  }
  // K entries -> positions in L, and the rows of A they combine
  m->kpos.clear();
  m->ka.clear();
  m->kb.clear();
  for (int jn = 0; jn < nw; jn++)
    for (int in : Kcol[jn]) {
      const int *b = m->Li.data() + m->Lp[jn], *e = m->Li.data() + m->Lp[jn + 1];
      const int *at = std::lower_bound(b, e, in);
      if (at == e || *at != in) return 1;
      m->kpos.push_back((int)(at - m->Li.data()));
      m->ka.push_back(m->perm[in]);
      m->kb.push_back(m->perm[jn]);
    }
  // level sets of the elimination tree
  std::vector<int> level(nw, 0);
  int nlev = nw > 0 ? 1 : 0;
  for (int j = 0; j < nw; j++) {
    for (int c : children[j]) level[j] = std::max(level[j], level[c] + 1);
    nlev = std::max(nlev, level[j] + 1);
  }
  m->level_ptr.assign(nlev + 1, 0);
  for (int j = 0; j < nw; j++) m->level_ptr[level[j] + 1]++;
  for (int l = 0; l < nlev; l++) m->level_ptr[l + 1] += m->level_ptr[l];
  m->level_cols.resize(nw);
  {
    std::vector<int> next(m->level_ptr.begin(), m->level_ptr.end() - 1);
    for (int j = 0; j < nw; j++) m->level_cols[next[level[j]]++] = j;
  }
  m->launches.clear();
  for (int l = 0; l < nlev;) {
    const int cnt = m->level_ptr[l + 1] - m->level_ptr[l];
    if (cnt == 1) {
      int l1 = l + 1;
      while (l1 < nlev && m->level_ptr[l1 + 1] - m->level_ptr[l1] == 1) l1++;
      m->launches.push_back({l, l1, 1, 1});
      l = l1;
    } else {
      m->launches.push_back({l, l + 1, cnt, 0});
      l++;
    }
  }
  return 0;
}

// ------------------------------------------------------------------ kernels
#define SP_THREADS 128

// K = C + A D^-1 A^T on the pattern of K, written into the (zeroed) storage of L
__global__ void __launch_bounds__(SP_THREADS)
    sp_assemble_kernel(int nk, const int *__restrict__ kpos, const int *__restrict__ ka,
                       const int *__restrict__ kb, const int *__restrict__ rowp,
                       const int *__restrict__ scols, const int *__restrict__ sidx,
                       const double *__restrict__ data, const double *__restrict__ Dinv,
                       const double *__restrict__ C, double *__restrict__ Lx) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nk) return;
  const int a = ka[e], b = kb[e];
  int pa = rowp[a], pb = rowp[b];
  const int ea = rowp[a + 1], eb = rowp[b + 1];
  double v = 0.0;
  while (pa < ea && pb < eb) {
    const int ca = scols[pa], cb = scols[pb];
    if (ca == cb) {
      v = fma(data[sidx[pa]] * Dinv[ca], data[sidx[pb]], v);
      pa++;
      pb++;
    } else if (ca < cb) {
      pa++;
    } else {
      pb++;
    }
  }
  if (a == b) v += C[a];
  Lx[kpos[e]] = v;
}

// Left-looking Cholesky of one column by one thread block.  `where` (one int per row of
// the matrix, private to the block) maps a row index to its position in column j: the rows
// of every contributing column's tail are rows of column j (fill property), so only
// entries written for this column are ever read.
__device__ void sp_chol_column(int j, const int *__restrict__ Lp, const int *__restrict__ Li,
                               const int *__restrict__ Rp, const int *__restrict__ Rk,
                               const int *__restrict__ Rpos, double *Lx, const int *perm,
                               int *fail, int *where) {
  const int c0 = Lp[j], c1 = Lp[j + 1];
  for (int t = c0 + threadIdx.x; t < c1; t += blockDim.x) where[Li[t]] = t;
  __syncthreads();
  for (int r = Rp[j]; r < Rp[j + 1]; r++) {  // fixed order: columns ascending
    const int k = Rk[r], pos = Rpos[r];
    const double ljk = Lx[pos];
    const int kend = Lp[k + 1];
    for (int t = pos + threadIdx.x; t < kend; t += blockDim.x) Lx[where[Li[t]]] -= Lx[t] * ljk;
    __syncthreads();
  }
  double d = Lx[c0];
  __syncthreads();
  if (!(d > 0.0)) {
    if (threadIdx.x == 0) atomicMin(fail, perm[j] + 1);
    d = 1.0;  // keep going with a harmless pivot; the caller sees the failure
  }
  const double s = sqrt(d);
  for (int t = c0 + threadIdx.x; t < c1; t += blockDim.x) Lx[t] = (t == c0) ? s : Lx[t] / s;
  __syncthreads();
}

__global__ void __launch_bounds__(SP_THREADS)
    sp_chol_kernel(int l0, int l1, const int *__restrict__ level_ptr,
                   const int *__restrict__ level_cols, const int *__restrict__ Lp,
                   const int *__restrict__ Li, const int *__restrict__ Rp,
                   const int *__restrict__ Rk, const int *__restrict__ Rpos, double *Lx,
                   const int *__restrict__ perm, int *fail, int *where_all, int nw) {
  int *where = where_all + (size_t)blockIdx.x * nw;
  for (int l = l0; l < l1; l++) {  // more than one level only when each has a single column
    for (int idx = level_ptr[l] + blockIdx.x; idx < level_ptr[l + 1]; idx += gridDim.x)
      sp_chol_column(level_cols[idx], Lp, Li, Rp, Rk, Rpos, Lx, perm, fail, where);
    __threadfence_block();
  }
}

__device__ __forceinline__ double sp_warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// forward substitution L y = b, one warp per row (rows of a level are independent)
__global__ void __launch_bounds__(SP_THREADS)
    sp_forward_kernel(int l0, int l1, const int *__restrict__ level_ptr,
                      const int *__restrict__ level_cols, const int *__restrict__ Lp,
                      const int *__restrict__ Rp, const int *__restrict__ Rk,
                      const int *__restrict__ Rpos, const double *__restrict__ Lx, double *y) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int l = l0; l < l1; l++) {
    const int idx = level_ptr[l] + w;
    if (idx < level_ptr[l + 1]) {
      const int j = level_cols[idx];
      double s = 0.0;
      for (int r = Rp[j] + lane; r < Rp[j + 1]; r += 32) s = fma(Lx[Rpos[r]], y[Rk[r]], s);
      s = sp_warp_sum(s);
      if (lane == 0) y[j] = (y[j] - s) / Lx[Lp[j]];
    }
    __syncwarp();
    __threadfence_block();
  }
}

// backward substitution L^T x = y, one warp per column, levels in reverse
__global__ void __launch_bounds__(SP_THREADS)
    sp_backward_kernel(int l0, int l1, const int *__restrict__ level_ptr,
                       const int *__restrict__ level_cols, const int *__restrict__ Lp,
                       const int *__restrict__ Li, const double *__restrict__ Lx, double *x) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (int l = l1 - 1; l >= l0; l--) {
    const int idx = level_ptr[l] + w;
    if (idx < level_ptr[l + 1]) {
      const int j = level_cols[idx];
      double s = 0.0;
      for (int t = Lp[j] + 1 + lane; t < Lp[j + 1]; t += 32) s = fma(Lx[t], x[Li[t]], s);
      s = sp_warp_sum(s);
      if (lane == 0) x[j] = (x[j] - s) / Lx[Lp[j]];
    }
    __syncwarp();
    __threadfence_block();
  }
}

// work[jn] = bw[perm[jn]] - (A D^-1 bx)[perm[jn]]   (bw may be null)
__global__ void __launch_bounds__(SP_THREADS)
    sp_rhs_kernel(int nw, const int *__restrict__ perm, const int *__restrict__ rowp,
                  const int *__restrict__ cols, const double *__restrict__ data,
                  const double *__restrict__ Dinv, const double *__restrict__ bx,
                  const double *__restrict__ bw, double *__restrict__ work) {
  const int jn = blockIdx.x * blockDim.x + threadIdx.x;
  if (jn >= nw) return;
  const int i = perm[jn];
  double s = 0.0;
  for (int e = rowp[i]; e < rowp[i + 1]; e++) {
    const int c = cols[e];
    s = fma(data[e], Dinv[c] * bx[c], s);
  }
  work[jn] = (bw ? bw[i] : 0.0) - s;
}

__global__ void __launch_bounds__(SP_THREADS)
    sp_scatter_kernel(int nw, const int *__restrict__ perm, const double *__restrict__ work,
                      double *__restrict__ yw) {
  const int jn = blockIdx.x * blockDim.x + threadIdx.x;
  if (jn < nw) yw[perm[jn]] = work[jn];
}

// yx = D^-1 (bx + A^T yw)
__global__ void __launch_bounds__(SP_THREADS)
    sp_finish_kernel(int nv, const int *__restrict__ colp, const int *__restrict__ rows,
                     const int *__restrict__ tmap, const double *__restrict__ data,
                     const double *__restrict__ Dinv, const double *__restrict__ bx,
                     const double *__restrict__ yw, double *__restrict__ yx) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nv) return;
  double s = 0.0;
  for (int f = colp[k]; f < colp[k + 1]; f++) s = fma(data[tmap[f]], yw[rows[f]], s);
  yx[k] = Dinv[k] * (bx[k] + s);
}

// out += alpha A px   /   out += alpha A^T pzw   (ParOptProblem.cpp:756-816)
__global__ void __launch_bounds__(SP_THREADS)
    sp_mult_kernel(int nw, const int *__restrict__ rowp, const int *__restrict__ cols,
                   const double *__restrict__ data, double alpha,
                   const double *__restrict__ px, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nw) return;
  double s = 0.0;
  for (int e = rowp[i]; e < rowp[i + 1]; e++) s = fma(data[e], px[cols[e]], s);
  out[i] += alpha * s;
}
__global__ void __launch_bounds__(SP_THREADS)
    sp_mult_t_kernel(int nv, const int *__restrict__ colp, const int *__restrict__ rows,
                     const int *__restrict__ tmap, const double *__restrict__ data,
                     double alpha, const double *__restrict__ pzw, double *__restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nv) return;
  double s = 0.0;
  for (int f = colp[k]; f < colp[k + 1]; f++) s = fma(data[tmap[f]], pzw[rows[f]], s);
  out[k] += alpha * s;
}

template <class T>
int upload(T **dst, const std::vector<T> &src) {
  const size_t bytes = sizeof(T) * (src.size() > 0 ? src.size() : 1);
  if (cudaMalloc((void **)dst, bytes) != cudaSuccess) return 1;
  if (!src.empty() &&
      cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice) != cudaSuccess)
    return 1;
  return 0;
}

int grid_for(int n) { return n > 0 ? (n + SP_THREADS - 1) / SP_THREADS : 1; }

// in place on d_work (permuted numbering)
int solve_permuted(pcu_sparsemat *m) {
  cudaStream_t st = m->ctx->stream;
  for (size_t a = 0; a < m->launches.size(); a++) {
    const SparseLaunch &L = m->launches[a];
    const int grid = L.serial ? 1 : (L.grid * 32 + SP_THREADS - 1) / SP_THREADS;
    sp_forward_kernel<<<grid, L.serial ? 32 : SP_THREADS, 0, st>>>(
        L.l0, L.l1, m->d_level_ptr, m->d_level_cols, m->d_Lp, m->d_Rp, m->d_Rk, m->d_Rpos,
        m->d_Lx, m->d_work);
    m->ctx->launches++;
  }
  for (size_t a = m->launches.size(); a-- > 0;) {
    const SparseLaunch &L = m->launches[a];
    const int grid = L.serial ? 1 : (L.grid * 32 + SP_THREADS - 1) / SP_THREADS;
    sp_backward_kernel<<<grid, L.serial ? 32 : SP_THREADS, 0, st>>>(
        L.l0, L.l1, m->d_level_ptr, m->d_level_cols, m->d_Lp, m->d_Li, m->d_Lx, m->d_work);
    m->ctx->launches++;
  }
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

int apply(pcu_sparsemat *m, pcu_vec *bx, pcu_vec *bw, pcu_vec *yx, pcu_vec *yw) {
  if (!m || !m->ctx || !m->factored || !bx || !yx || !yw) return 1;
  if (bx->n != m->nvars || yx->n != m->nvars || yw->n != m->nwcon || (bw && bw->n != m->nwcon))
    return 1;
  pcu_vec_ready(bx);
  pcu_vec_ready(bw);
  pcu_vec_ready(yx);
  pcu_vec_ready(yw);
  pcu_vec_ready(m->Dinv);
  cudaStream_t st = m->ctx->stream;
  if (m->nwcon > 0) {
    sp_rhs_kernel<<<grid_for(m->nwcon), SP_THREADS, 0, st>>>(
        m->nwcon, m->d_perm, m->d_rowp, m->d_cols, m->d_data, m->Dinv->d, bx->d,
        bw ? bw->d : nullptr, m->d_work);
    m->ctx->launches++;
    if (solve_permuted(m)) return 1;
    sp_scatter_kernel<<<grid_for(m->nwcon), SP_THREADS, 0, st>>>(m->nwcon, m->d_perm, m->d_work,
                                                                 yw->d);
    m->ctx->launches++;
  }
  if (m->nvars > 0) {
    sp_finish_kernel<<<grid_for(m->nvars), SP_THREADS, 0, st>>>(
        m->nvars, m->d_colp, m->d_rows, m->d_tmap, m->d_data, m->Dinv->d, bx->d, yw->d, yx->d);
    m->ctx->launches++;
  }
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

pcu_sparsemat *pcu_sparsemat_create(pcu_ctx *ctx, int nvars, int nwcon, const int *rowp,
                                    const int *cols, int ordering) {
  if (nvars < 0 || nwcon < 0 || !rowp || (rowp[nwcon] > 0 && !cols) || rowp[0] != 0) return nullptr;
  for (int i = 0; i < nwcon; i++)
    if (rowp[i + 1] < rowp[i]) return nullptr;
  pcu_sparsemat *m = new pcu_sparsemat;
  m->ctx = ctx;
  m->nvars = nvars;
  m->nwcon = nwcon;
  m->rowp.assign(rowp, rowp + nwcon + 1);
  m->cols.assign(cols, cols + rowp[nwcon]);
  const int src = symbolic(m, ordering);
  if (src) {
    fprintf(stderr, src == 2 ? "paropt_b200: pcu_sparsemat_create: the Cholesky factor exceeds 4e8 entries "
                               "with this ordering\n"
                             : "paropt_b200: pcu_sparsemat_create: column index out of range or "
                               "duplicate entry\n");
    delete m;
    return nullptr;
  }
  if (!ctx) return m;  // symbolic data only (CPU tests of the host phase)
  int bad = 0;
  bad |= upload(&m->d_rowp, m->rowp) | upload(&m->d_cols, m->cols);
  bad |= upload(&m->d_srt_cols, m->srt_cols) | upload(&m->d_srt_idx, m->srt_idx);
  bad |= upload(&m->d_colp, m->colp) | upload(&m->d_rows, m->rows) | upload(&m->d_tmap, m->tmap);
  bad |= upload(&m->d_perm, m->perm) | upload(&m->d_Lp, m->Lp) | upload(&m->d_Li, m->Li);
  bad |= upload(&m->d_Rp, m->Rp) | upload(&m->d_Rk, m->Rk) | upload(&m->d_Rpos, m->Rpos);
  bad |= upload(&m->d_kpos, m->kpos) | upload(&m->d_ka, m->ka) | upload(&m->d_kb, m->kb);
  bad |= upload(&m->d_level_ptr, m->level_ptr) | upload(&m->d_level_cols, m->level_cols);
  const size_t nl = m->Li.size() > 0 ? m->Li.size() : 1;
  bad |= cudaMalloc((void **)&m->d_Lx, sizeof(double) * nl) != cudaSuccess;
  bad |= cudaMalloc((void **)&m->d_data, sizeof(double) * (m->nnz > 0 ? m->nnz : 1)) != cudaSuccess;
  bad |= cudaMalloc((void **)&m->d_work, sizeof(double) * (nwcon > 0 ? nwcon : 1)) != cudaSuccess;
  bad |= cudaMalloc((void **)&m->d_fail, sizeof(int)) != cudaSuccess;
  {
    // concurrent columns of the factorisation: two blocks per SM, at most 256 MB of maps
    int widest = 1;
    for (const SparseLaunch &L : m->launches) widest = std::max(widest, L.grid);
    long long slots = std::min<long long>(widest, 2LL * ctx->num_sms);
    const long long cap = (256LL << 20) / (4LL * std::max(nwcon, 1));
    slots = std::max<long long>(1, std::min(slots, cap));
    m->nslots = (int)slots;
    bad |= cudaMalloc((void **)&m->d_where, sizeof(int) * (size_t)slots * (size_t)std::max(nwcon, 1)) !=
           cudaSuccess;
  }
  if (!bad) bad |= cudaMemset(m->d_data, 0, sizeof(double) * (m->nnz > 0 ? m->nnz : 1)) != cudaSuccess;
  if (bad) {
    fprintf(stderr, "paropt_b200: pcu_sparsemat_create: device allocation failed\n");
    pcu_sparsemat_destroy(m);
    return nullptr;
  }
  return m;
}

void pcu_sparsemat_destroy(pcu_sparsemat *m) {
  if (!m) return;
  if (m->ctx) cudaStreamSynchronize(m->ctx->stream);
  void *ptrs[] = {m->d_rowp, m->d_cols, m->d_srt_cols, m->d_srt_idx, m->d_colp, m->d_rows,
                  m->d_tmap, m->d_perm, m->d_Lp, m->d_Li, m->d_Rp, m->d_Rk, m->d_Rpos,
                  m->d_kpos, m->d_ka, m->d_kb, m->d_level_ptr, m->d_level_cols, m->d_fail, m->d_where,
                  m->d_data, m->d_Lx, m->d_work};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  delete m;
}

int pcu_sparsemat_set_data(pcu_sparsemat *m, const double *data) {
  if (!m || !m->ctx || !data) return 1;
  if (m->nnz > 0) {
    PCU_CUDA_OK(cudaMemcpyAsync(m->d_data, data, sizeof(double) * (size_t)m->nnz,
                                cudaMemcpyHostToDevice, m->ctx->stream));
    // the caller may change `data` as soon as this returns (pageable source)
    PCU_CUDA_OK(cudaStreamSynchronize(m->ctx->stream));
  }
  return 0;
}

double *pcu_sparsemat_data_device_ptr(pcu_sparsemat *m) { return m ? m->d_data : nullptr; }

int pcu_sparsemat_factor(pcu_sparsemat *m, pcu_vec *x, pcu_vec *Dinv, pcu_vec *C) {
  (void)x;
  if (!m || !m->ctx || !Dinv || !C) return -1;
  if (Dinv->n != m->nvars || C->n != m->nwcon) return -1;
  pcu_vec_ready(Dinv);
  pcu_vec_ready(C);
  m->Dinv = Dinv;
  cudaStream_t st = m->ctx->stream;
  const int big = 0x7fffffff;
  if (cudaMemcpyAsync(m->d_fail, &big, sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess)
    return -1;
  if (cudaMemsetAsync(m->d_Lx, 0, sizeof(double) * (m->Li.size() > 0 ? m->Li.size() : 1), st) !=
      cudaSuccess)
    return -1;
  const int nk = (int)m->kpos.size();
  if (nk > 0) {
    sp_assemble_kernel<<<grid_for(nk), SP_THREADS, 0, st>>>(
        nk, m->d_kpos, m->d_ka, m->d_kb, m->d_rowp, m->d_srt_cols, m->d_srt_idx, m->d_data,
        Dinv->d, C->d, m->d_Lx);
    m->ctx->launches++;
  }
  for (const SparseLaunch &L : m->launches) {
    const int grid = L.grid < m->nslots ? L.grid : m->nslots;
    sp_chol_kernel<<<grid, SP_THREADS, 0, st>>>(L.l0, L.l1, m->d_level_ptr, m->d_level_cols,
                                                m->d_Lp, m->d_Li, m->d_Rp, m->d_Rk, m->d_Rpos,
                                                m->d_Lx, m->d_perm, m->d_fail, m->d_where,
                                                m->nwcon);
    m->ctx->launches++;
  }
  int fail = 0;
  if (cudaMemcpyAsync(&fail, m->d_fail, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess)
    return -1;
  m->factored = 1;
  return fail == big ? 0 : fail;
}

int pcu_sparsemat_apply3(pcu_sparsemat *m, pcu_vec *bx, pcu_vec *yx, pcu_vec *yw) {
  return apply(m, bx, nullptr, yx, yw);
}
int pcu_sparsemat_apply4(pcu_sparsemat *m, pcu_vec *bx, pcu_vec *bw, pcu_vec *yx, pcu_vec *yw) {
  if (!bw) return 1;
  return apply(m, bx, bw, yx, yw);
}

int pcu_sparsemat_mult_add(pcu_sparsemat *m, double alpha, pcu_vec *px, pcu_vec *out) {
  if (!m || !m->ctx || !px || !out || px->n != m->nvars || out->n != m->nwcon) return 1;
  pcu_vec_ready(px);
  pcu_vec_ready(out);
  if (m->nwcon > 0) {
    sp_mult_kernel<<<grid_for(m->nwcon), SP_THREADS, 0, m->ctx->stream>>>(
        m->nwcon, m->d_rowp, m->d_cols, m->d_data, alpha, px->d, out->d);
    m->ctx->launches++;
  }
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

int pcu_sparsemat_mult_transpose_add(pcu_sparsemat *m, double alpha, pcu_vec *pzw, pcu_vec *out) {
  if (!m || !m->ctx || !pzw || !out || pzw->n != m->nwcon || out->n != m->nvars) return 1;
  pcu_vec_ready(pzw);
  pcu_vec_ready(out);
  if (m->nvars > 0) {
    sp_mult_t_kernel<<<grid_for(m->nvars), SP_THREADS, 0, m->ctx->stream>>>(
        m->nvars, m->d_colp, m->d_rows, m->d_tmap, m->d_data, alpha, pzw->d, out->d);
    m->ctx->launches++;
  }
  PCU_CUDA_OK(cudaGetLastError());
  return 0;
}

int pcu_sparsemat_info(pcu_sparsemat *m, int *nnzK, int *nnzL, int *nlevels, int *nlaunches) {
  if (!m) return 1;
  if (nnzK) *nnzK = m->nnzK;
  if (nnzL) *nnzL = (int)m->Li.size();
  if (nlevels) *nlevels = (int)m->level_ptr.size() - 1;
  if (nlaunches) *nlaunches = (int)m->launches.size();
  return 0;
}

// The symbolic factorisation (host arrays; CPU tests emulate the numeric phase on them).
// Sizes: perm nwcon; Lp nwcon + 1; Li nnzL; Rp nwcon + 1; Rk, Rpos nnzL - nwcon;
// kpos, ka, kb nnzK; level_ptr nlevels + 1; level_cols nwcon.  Null pointers are skipped.
int pcu_sparsemat_symbolic(pcu_sparsemat *m, int *perm, int *Lp, int *Li, int *Rp, int *Rk,
                           int *Rpos, int *kpos, int *ka, int *kb, int *level_ptr,
                           int *level_cols) {
  if (!m) return 1;
  auto put = [](int *dst, const std::vector<int> &src) {
    if (dst && !src.empty()) memcpy(dst, src.data(), sizeof(int) * src.size());
  };
  put(perm, m->perm);
  put(Lp, m->Lp);
  put(Li, m->Li);
  put(Rp, m->Rp);
  put(Rk, m->Rk);
  put(Rpos, m->Rpos);
  put(kpos, m->kpos);
  put(ka, m->ka);
  put(kb, m->kb);
  put(level_ptr, m->level_ptr);
  put(level_cols, m->level_cols);
  return 0;
}

}  // extern "C"
