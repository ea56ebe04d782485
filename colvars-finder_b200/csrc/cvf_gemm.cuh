// cvf_gemm.cuh -- descriptor of the dense fp32 products of the layer-wise autoencoder path (cvf_ae_wide.cu), shared by the
// SIMT kernel (FFMA2, cvf_ae_wide.cu) and the tensor-core kernel (tcgen05 / TMEM, cvf_gemm_tc.cu).
#pragma once
#include <cuda_runtime.h>

namespace cvf {
namespace wide {

enum Epilogue { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_TANH = 2, EPI_MUL_OM = 3, EPI_BIAS_LOSS = 4 };

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
  // tensor-core path: the same activations as their K-major image (rows = m, K = n; hi + lo restores the fp32 value exactly),
  // so that the forward products need not store a row-major copy; used instead of `act` when not NULL
  const float* act_img;
  int act_img_kblocks;
  // Tensor-core path only: operand "tile images".  An image holds the operand already split into TF32 hi / lo parts and laid
  // out exactly as the kernel's shared-memory stage wants it (K-major 128 x 32 tiles, 128-byte swizzle, zero padded):
  // image[(row_tile * kblocks + k_block) * 8192 floats] = hi tile (4096 floats) | lo tile, k_block counted from K = 0.  A stage
  // is then filled by one 1-D bulk asynchronous copy (TMA) per operand instead of through registers.  NULL = stage from A / B.
  const float* a_img;
  const float* b_img;
  int a_img_kblocks, b_img_kblocks;
  // Tensor-core path only, single split: the epilogue also writes C as operand images for the products that consume it, so
  // that no product stages an operand through registers and no separate pass re-reads C to build an image:
  //   c_img_k  C as an operand with rows = m and K = n (the A operand of the next layer's product); k-blocks = ceil(N / 32)
  //   c_img_t  C^T as an operand with rows = n and K = m (an operand of a weight-gradient product, whose K index is the
  //            frame); k-blocks = ceil(M / 32).  c_img_t_ones > 0 adds a row of ones (for m < M) at that row index (= N), which
  //            turns the bias gradient into one more column of the weight-gradient product.
  // Entries outside the matrix are written as zeros.  C itself is not stored when C is NULL.
  float* c_img_k;
  float* c_img_t;
  int c_img_k_kblocks, c_img_t_kblocks, c_img_t_ones;
  // EPI_BIAS_LOSS (tensor-core path, last layer of the autoencoder): with out = product + bias, e = out - loss_in[m][n],
  // C / the images receive delta = 2 w[m] e, and every output tile t (column tile fastest) writes its share of
  // (sum w |e|^2, sum w) to loss_part[2 t], loss_part[2 t + 1]
  const float* loss_in;
  long long loss_ld;
  const float* loss_w;
  double* loss_part;
};

// cvf_gemm_tc.cu: the same product on the 5th-generation tensor cores; both operands as tile images; ceil(N/128) x ceil(M/128) x
// splits output tiles walked by one persistent CTA per SM; k_per_split a multiple of 32 when splits > 1
int launch_gemm_tc(const Gemm& g, int splits, cudaStream_t stream);
inline int gemm_tc_tiles(const Gemm& g, int splits) { return ((g.N + 127) / 128) * ((g.M + 127) / 128) * splits; }
// floats of the image of an operand with `rows` rows and K columns
inline size_t tile_image_floats(int rows, int K) { return (size_t)((rows + 127) / 128) * ((K + 31) / 32) * 8192; }
// builds the image of Xop[row][k] = kcontig ? X[row * ld + k] : X[k * ld + row], row < rows; ones_row > 0 (>= rows): one more
// row, all ones for k < K
int launch_tile_image(const float* X, long long ld, int kcontig, int rows, int K, int ones_row, float* img, cudaStream_t stream);

}  // namespace wide
}  // namespace cvf
