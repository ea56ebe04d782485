// cvf_gemm.cuh -- descriptor of the dense fp32 products of the layer-wise autoencoder path (cvf_ae_wide.cu), shared by the
// SIMT kernel (FFMA2, cvf_ae_wide.cu) and the tensor-core kernel (tcgen05 / TMEM, cvf_gemm_tc.cu).
#pragma once
#include <cuda_runtime.h>

namespace cvf {
namespace wide {

enum Epilogue { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_TANH = 2, EPI_MUL_OM = 3 };

struct Gemm {
  // C[m][n] = sum_k Aop[m][k] Bop[k][n],  m < M, n < N, k in this split's range.  Every matrix has a leading dimension that
  // is a multiple of 4 floats and a 16-byte aligned base.
  const float* A;
  long long lda;
  int a_kcontig;   // 1: Aop[m][k] = A[m * lda + k];   0: Aop[m][k] = A[k * lda + m]
  const float* B;
  long long ldb;
  int b_kcontig;   // 1: Bop[k][n] = B[n * ldb + k];   0: Bop[k][n] = B[k * ldb + n]
  float* C;
  long long ldc;
  long long c_split_stride;   // floats between the outputs of consecutive k-splits (0: single split)
  int M, N, K, k_per_split;
  int epi;
  const float* bias;   // [N]
  const float* act;    // EPI_MUL_OM: [M][ldc] activations A with C *= 1 - A^2
};

// cvf_gemm_tc.cu: the same product on the 5th-generation tensor cores; grid = (ceil(N/128), ceil(M/128), splits), k_per_split
// a multiple of 32 when splits > 1
int launch_gemm_tc(const Gemm& g, int splits, cudaStream_t stream);

}  // namespace wide
}  // namespace cvf
