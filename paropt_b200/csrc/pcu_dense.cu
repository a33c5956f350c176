// pcu_dense.cu -- device side of pcu_dense.cuh: a single-CTA kernel that runs the
// small dense algebra of the KKT solve between two streaming passes, so that the
// iteration's chain  Gram -> pass 2 + residual -> pass 2 + statistics  needs no host
// round trip (LU of G and Ce, SMW coefficient assembly, dense residuals, fed by the
// in-stream all-reduce / all-gather result).
//
// The flat work buffer (<= 5k doubles) is staged in shared memory by all threads, warp
// 0 runs the lane-parallel algorithm of pcu_dense.cuh on it (matrices of at most
// 32 x 32), all threads write the result region back.
//
// Built with -fmad=false (paropt_b200/build.py): identical multiply / add sequence to
// the host compiler's, so chain and host path agree bit for bit on the same inputs.
#include <stdio.h>

#include "pcu_dense.cuh"

#define PCU_DENSE_THREADS 128

__global__ void __launch_bounds__(PCU_DENSE_THREADS, 1)
    pcu_dense_kernel(double *buf, const DenseOff o, const int phase, const double *Sin,
                     const double *red, const int world, const int stride) {
  extern __shared__ double2 dense_smem2[];
  double *w = reinterpret_cast<double *>(dense_smem2);
  double *scratch = w + o.total;
  const int tid = threadIdx.x;
  // phase 0 needs the inputs only (the rest is produced here); phase 1 everything
  const int n_in = phase == 0 ? o.S : o.total;
  for (int i = tid; i < n_in; i += PCU_DENSE_THREADS) w[i] = buf[i];
  if (phase == 0) {
    const int nS = o.ld * o.ld;
    for (int i = tid; i < nS; i += PCU_DENSE_THREADS) w[o.S + i] = Sin[i];
  }
  __syncthreads();
  if (tid < 32) {
    if (phase == 0) pcu_dense_phase_a(w, o, scratch);
    else pcu_dense_phase_b(w, o, red, world, stride, scratch);
  }
  __syncthreads();
  for (int i = o.S + tid; i < o.total; i += PCU_DENSE_THREADS) buf[i] = w[i];
}

// Enqueues one phase on `stream`.  buf: device work buffer (layout `o`); Sin: Gram
// result (phase 0); red / world / stride: rank-ordered partial reductions (phase 1).
int pcu_dense_enqueue(cudaStream_t stream, double *buf, const DenseOff &o, int phase,
                      const double *Sin, const double *red, int world, int stride) {
  const size_t smem = sizeof(double) * (size_t)(o.total + PCU_DENSE_SCRATCH(o.c, o.q));
  if (smem > 48 * 1024) {
    static bool raised[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !raised[dev]) {
      if (cudaFuncSetAttribute(pcu_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               96 * 1024) != cudaSuccess)
        return 1;
      if (dev >= 0 && dev < 64) raised[dev] = true;
    }
    if (smem > 96 * 1024) return 1;
  }
  pcu_dense_kernel<<<1, PCU_DENSE_THREADS, smem, stream>>>(buf, o, phase, Sin, red, world, stride);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// The same two phases on the host (checker of the device path): identical statements,
// identical rounding.
void pcu_dense_host(double *w, const DenseOff &o, int phase, const double *Sin,
                    const double *red, int world, int stride) {
  double scratch[PCU_DENSE_SCRATCH(PCU_DENSE_MAXM / 2, PCU_DENSE_MAXM / 2) + 8 * PCU_DENSE_MAXM];
  if (phase == 0) {
    for (int i = 0; i < o.ld * o.ld; i++) w[o.S + i] = Sin[i];
    pcu_dense_phase_a(w, o, scratch);
  } else {
    pcu_dense_phase_b(w, o, red, world, stride, scratch);
  }
}
