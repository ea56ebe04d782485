// cvf_ae_wide.cu -- AutoEncoderTask.weighted_MSE_loss + backward (core.py:652-666,708) for networks whose weights do not
// fit shared memory (e.g. [3000,512,512,2] + [2,512,512,3000]): layer by layer, every layer a dense fp32 product over a
// chunk of frames.
//
//   forward   A_l  = act(A_{l-1} W_l^T + b_l)                 C[M x N] = A[M x K] B[N x K]^T     epilogue: + bias, tanh
//   loss      e = A_L - A_0, sums (sum w |e|^2, sum w), delta_L = 2 w e
//   backward  delta_{l-1} = (delta_l W_l) .* (1 - A_{l-1}^2)  C[M x K] = A[M x N] B[N x K]       epilogue: * (1 - act^2)
//   gradient  [dW_l | db_l] = delta_l^T [A_{l-1} | 1]         C[N x (K+1)] = sum over frames, split over the frames
//
// Two implementations of the three products (cvf_ae_set_wide_path):
//   * tensor cores (default): cvf_gemm_tc.cu, tcgen05 with the 3 x TF32 split.  Every operand of every product is a "tile image"
//     (cvf_gemm.cuh) filled by one bulk copy per stage: the weights' images are built once per step, and every activation /
//     delta is written as images by the epilogue of the product that computes it (K-major for the next layer's product,
//     transposed -- K = frame -- for the weight-gradient product), so no product stages an operand through registers and no
//     separate pass re-reads an activation.  Only the input features and delta_L (written by the loss kernel) go through
//     tile_image_kernel.
//   * fp32 SIMT (mode 1, kept as the cross-check of the tests): one SGEMM kernel, 128 x 128 x 16 CTA tiles, 8 x 8 register tiles
//     with packed FFMA2, operands staged through shared memory k-major, global loads of tile i+1 in flight during tile i.
//
// Every matrix the products touch lives in the caller's workspace with a leading dimension padded to a multiple of 4 floats
// and zero padding, so all global accesses are aligned 128-bit.  Activation buffers carry one extra column of ones, which
// turns the bias gradient into the last column of the weight-gradient product.
#include <atomic>
#include <string.h>

#include "cvf_common.cuh"
#include "cvf_gemm.cuh"

namespace cvf {
namespace wide {

typedef unsigned long long u64;
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)), "l"(*reinterpret_cast<u64*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}

constexpr int BM = 128, BN = 128, BK = 16, LDS_ = BM + 4;   // shared tiles are [BK][BM + 4] floats

// Load one BK x 128 operand tile into registers (2 float4 per thread), then store it k-major into shared memory.
struct TileRegs {
  float4 v[2];
};

__device__ __forceinline__ void load_tile(TileRegs& r, const float* __restrict__ P, long long ld, int kcontig, int row0, int nrows,
                                          int k0, int k1, int tid) {
  // rows = the m (or n) index of the tile, bounded by nrows; k bounded by k1.  Padding is zero-filled.
  if (kcontig) {
    // element (row, k): P[row * ld + k]; thread -> (row = idx / 4, 4 consecutive k)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + 256 * i, row = row0 + (idx >> 2), k = k0 + 4 * (idx & 3);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < nrows && k < k1) {
        v = __ldg(reinterpret_cast<const float4*>(P + (size_t)row * ld + k));
        if (k + 3 >= k1) {   // the split's range ends inside this vector
          if (k + 1 >= k1) v.y = 0.f;
          if (k + 2 >= k1) v.z = 0.f;
          v.w = 0.f;
        }
      }
      r.v[i] = v;
    }
  } else {
    // element (row, k): P[k * ld + row]; thread -> (k = idx / 32, 4 consecutive rows)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + 256 * i, k = k0 + (idx >> 5), row = row0 + 4 * (idx & 31);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < k1 && row < nrows) {
        v = __ldg(reinterpret_cast<const float4*>(P + (size_t)k * ld + row));
        if (row + 3 >= nrows) {
          if (row + 1 >= nrows) v.y = 0.f;
          if (row + 2 >= nrows) v.z = 0.f;
          v.w = 0.f;
        }
      }
      r.v[i] = v;
    }
  }
}

__device__ __forceinline__ void store_tile(float* __restrict__ S, const TileRegs& r, int kcontig, int tid) {
  if (kcontig) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + 256 * i, row = idx >> 2, k = 4 * (idx & 3);
      S[(k + 0) * LDS_ + row] = r.v[i].x;
      S[(k + 1) * LDS_ + row] = r.v[i].y;
      S[(k + 2) * LDS_ + row] = r.v[i].z;
      S[(k + 3) * LDS_ + row] = r.v[i].w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + 256 * i, k = idx >> 5, row = 4 * (idx & 31);
      *reinterpret_cast<float4*>(S + k * LDS_ + row) = r.v[i];
    }
  }
}

__device__ __forceinline__ float tanh_ref(float x) { return cvf_tanh(x); }

__global__ void __launch_bounds__(256, 2) sgemm_kernel(const Gemm g) {
  __shared__ __align__(16) float As[2][BK * LDS_];
  __shared__ __align__(16) float Bs[2][BK * LDS_];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * g.k_per_split, kend = min(g.K, kbeg + g.k_per_split);
  float2 acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  TileRegs ra, rb;
  load_tile(ra, g.A, g.lda, g.a_kcontig, m0, g.M, kbeg, kend, tid);
  load_tile(rb, g.B, g.ldb, g.b_kcontig, n0, g.N, kbeg, kend, tid);
  store_tile(As[0], ra, g.a_kcontig, tid);
  store_tile(Bs[0], rb, g.b_kcontig, tid);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) {
      load_tile(ra, g.A, g.lda, g.a_kcontig, m0, g.M, k0 + BK, kend, tid);
      load_tile(rb, g.B, g.ldb, g.b_kcontig, n0, g.N, k0 + BK, kend, tid);
    }
    const float* a = As[buf];
    const float* b = Bs[buf];
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(a + kk * LDS_ + 4 * ty);
      const float4 a1 = *reinterpret_cast<const float4*>(a + kk * LDS_ + 64 + 4 * ty);
      const float4 b0 = *reinterpret_cast<const float4*>(b + kk * LDS_ + 4 * tx);
      const float4 b1 = *reinterpret_cast<const float4*>(b + kk * LDS_ + 64 + 4 * tx);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float2 bv[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 ai = make_float2(av[i], av[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = ffma2(ai, bv[j], acc[i][j]);
      }
    }
    if (more) {
      store_tile(As[buf ^ 1], ra, g.a_kcontig, tid);
      store_tile(Bs[buf ^ 1], rb, g.b_kcontig, tid);
    }
    __syncthreads();
    buf ^= 1;
  }
  // epilogue: rows m0 + {4 ty .. 4 ty + 3, 64 + 4 ty ..}, columns n0 + {4 tx .., 64 + 4 tx ..}
  float* C = g.C + (size_t)blockIdx.z * g.c_split_stride;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + i - 4);
    if (m >= g.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + 64 * h + 4 * tx;
      if (n >= g.N) continue;
      float v[4] = {acc[i][2 * h].x, acc[i][2 * h].y, acc[i][2 * h + 1].x, acc[i][2 * h + 1].y};
      if (g.epi == EPI_BIAS || g.epi == EPI_BIAS_TANH) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (n + c < g.N) {
            v[c] += g.bias[n + c];
            if (g.epi == EPI_BIAS_TANH) v[c] = tanh_ref(v[c]);
          }
      } else if (g.epi == EPI_MUL_OM) {
        const float4 a4 = *reinterpret_cast<const float4*>(g.act + (size_t)m * g.ldc + n);
        v[0] *= fmaf(-a4.x, a4.x, 1.0f), v[1] *= fmaf(-a4.y, a4.y, 1.0f), v[2] *= fmaf(-a4.z, a4.z, 1.0f), v[3] *= fmaf(-a4.w, a4.w, 1.0f);
      }
      if (n + 3 < g.N) {
        *reinterpret_cast<float4*>(C + (size_t)m * g.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
      } else {   // the padding columns (and the ones column of an activation buffer) are not this kernel's to write
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (n + c < g.N) C[(size_t)m * g.ldc + n + c] = v[c];
      }
    }
  }
}

// rows [f0, f0 + M) of the caller's [B][d] features -> padded chunk buffer [M][ld] with the ones column at d
__global__ void __launch_bounds__(256) stage_input_kernel(const float* __restrict__ feat, long long f0, int M, int d, float* __restrict__ out,
                                                          int ld) {
  const long long n = (long long)M * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / ld), c = (int)(i - (long long)m * ld);
    out[i] = c < d ? __ldg(feat + (size_t)(f0 + m) * d + c) : (c == d ? 1.0f : 0.0f);
  }
}

// columns d .. ld-1 of an activation buffer [M][ld]: the ones column (bias gradient) and zero padding
__global__ void __launch_bounds__(256) set_pad_kernel(float* __restrict__ buf, int M, int d, int ld) {
  const int np = ld - d;
  const long long n = (long long)M * np;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / np), c = (int)(i - (long long)m * np);
    buf[(size_t)m * ld + d + c] = c == 0 ? 1.0f : 0.0f;
  }
}

// padded copies of the weights: W_l [d_out][ld_in] (zero padding), biases untouched
__global__ void __launch_bounds__(256) pad_weights_kernel(const float* __restrict__ W, int rows, int cols, float* __restrict__ out, int ld) {
  const int n = rows * ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / ld, c = i - r * ld;
    out[i] = c < cols ? W[(size_t)r * cols + c] : 0.0f;
  }
}

// e = out - in; delta = 2 w e (written over `out`); per-block partial sums of w |e|^2 and w
__global__ void __launch_bounds__(256) loss_delta_kernel(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ w,
                                                         long long f0, int M, int d, int ld, int ld_in, double* __restrict__ part) {
  __shared__ double red[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s2 = 0.0, s0 = 0.0;
  // a warp per frame row: coalesced along the feature dimension
  const int nwarps = gridDim.x * 8;
  for (int m = blockIdx.x * 8 + warp; m < M; m += nwarps) {
    const float wf = __ldg(w + f0 + m);
    float accf = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float e = out[(size_t)m * ld + c] - in[(size_t)m * ld_in + c];
      accf = fmaf(e, e, accf);
      out[(size_t)m * ld + c] = 2.0f * wf * e;
    }
    s2 += (double)wf * (double)accf;
    if (lane == 0) s0 += (double)wf;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o), s0 += __shfl_xor_sync(0xffffffffu, s0, o);
  if (lane == 0) red[0][warp] = s2, red[1][warp] = s0;
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int q = 0; q < 8; ++q) t += red[threadIdx.x][q];
    part[(size_t)blockIdx.x * 2 + threadIdx.x] = t;
  }
}

// grad[goff_w + r * cols + c] += sum_s P[s][r][c] (c < cols),  grad[goff_b + r] += sum_s P[s][r][cols]
__global__ void __launch_bounds__(256) accumulate_grad_kernel(const float* __restrict__ P, int splits, long long split_stride, int rows,
                                                              int cols, int ld, double* __restrict__ grad, int goff_w, int goff_b) {
  const int n = rows * (cols + 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / (cols + 1), c = i - r * (cols + 1);
    double s = 0.0;
    for (int q = 0; q < splits; ++q) s += (double)P[(size_t)q * split_stride + (size_t)r * ld + c];
    if (c < cols) grad[goff_w + (size_t)r * cols + c] += s;
    else grad[goff_b + r] += s;
  }
}

// sums[c] (+)= sum_b part[2 b + c], c = 0, 1: strided partial sums per thread, then a tree over the block (fixed order)
__global__ void __launch_bounds__(256) add_sums_kernel(const double* __restrict__ part, int n_blocks, double* __restrict__ sums, int first) {
  __shared__ double red[2][256];
  double t0 = 0.0, t1 = 0.0;
  for (int b = threadIdx.x; b < n_blocks; b += 256) t0 += part[(size_t)b * 2], t1 += part[(size_t)b * 2 + 1];
  red[0][threadIdx.x] = t0, red[1][threadIdx.x] = t1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[0][threadIdx.x] += red[0][threadIdx.x + o], red[1][threadIdx.x] += red[1][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 2) sums[threadIdx.x] = (first ? 0.0 : sums[threadIdx.x]) + red[threadIdx.x][0];
}

__global__ void zero_doubles_kernel(double* p, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = 0.0;
}

static inline int round4(int v) { return (v + 3) & ~3; }

struct WidePlan {
  int L;
  int dims[kMaxLayers + 1], ld[kMaxLayers + 1], act[kMaxLayers];
  int gw_off[kMaxLayers], gb_off[kMaxLayers];
  size_t w_off[kMaxLayers];   // padded weights (floats from the start of the weight area)
  size_t w_floats;
  size_t act_floats_per_frame;   // sum over layers 0..L of ld
  size_t delta_floats_per_frame; // 2 * max ld (ping-pong)
  size_t dw_max_floats;          // largest padded weight-gradient block
  // tensor-core path: tile images (cvf_gemm.cuh) of the weights as the B operand of the forward products (wf) and of the
  // back-propagation products (wb), and per frame the floats of the two operand images of a weight-gradient product
  size_t wf_off[kMaxLayers], wb_off[kMaxLayers];
  size_t wimg_floats;
  // per frame: K-major images of A_0 .. A_{L-1} (ak_off), transposed images with the ones row (at_off), and two ping-pong pairs
  // (K-major, transposed) for the deltas
  size_t ak_off[kMaxLayers], at_off[kMaxLayers], dk_floats, dt_floats;
  size_t img_floats_per_frame;
};

static void make_plan(const NetPlan& np, WidePlan* P) {
  P->L = np.L;
  size_t wf = 0;
  int maxld = 0;
  P->act_floats_per_frame = 0;
  P->dw_max_floats = 0;
  for (int l = 0; l <= np.L; ++l) {
    P->dims[l] = np.dims[l];
    P->ld[l] = round4(np.dims[l] + 1);   // room for the ones column
    P->act_floats_per_frame += P->ld[l];
    if (P->ld[l] > maxld) maxld = P->ld[l];
  }
  for (int l = 0; l < np.L; ++l) {
    P->act[l] = np.act[l];
    P->gw_off[l] = np.gw_off[l], P->gb_off[l] = np.gb_off[l];
    P->w_off[l] = wf;
    wf += (size_t)np.dims[l + 1] * P->ld[l];
    const size_t dw = (size_t)np.dims[l + 1] * P->ld[l];
    if (dw > P->dw_max_floats) P->dw_max_floats = dw;
  }
  P->w_floats = wf;
  P->delta_floats_per_frame = 2 * (size_t)maxld;
  size_t wi = 0;
  for (int l = 0; l < np.L; ++l) {
    P->wf_off[l] = wi, wi += tile_image_floats(np.dims[l + 1], np.dims[l]);
    P->wb_off[l] = wi, wi += tile_image_floats(np.dims[l], np.dims[l + 1]);
  }
  P->wimg_floats = wi;
  size_t fi = 0;
  int maxd = 0;
  for (int l = 0; l < np.L; ++l) {
    P->ak_off[l] = fi, fi += 2 * (size_t)((np.dims[l] + 31) / 32 * 32);          // hi + lo, K padded to the k-block
    P->at_off[l] = fi, fi += 2 * (size_t)((np.dims[l] + 1 + 127) / 128 * 128);   // hi + lo, rows (+ ones row) padded to the tile
    if (np.dims[l + 1] > maxd) maxd = np.dims[l + 1];
  }
  P->dk_floats = 2 * (size_t)((maxd + 31) / 32 * 32);
  P->dt_floats = 2 * (size_t)((maxd + 127) / 128 * 128);
  P->img_floats_per_frame = fi + 2 * (P->dk_floats + P->dt_floats);
}

constexpr int kMaxSplits = 32;
constexpr long long kChunkFrames = 32768;
// slots of (sum w |e|^2, sum w) partial sums: one per block of the loss kernel, or one per output tile of the last product
static size_t part_slots(const WidePlan& P) {
  const size_t tiles = (size_t)((P.dims[P.L] + 127) / 128) * (size_t)(kChunkFrames / 128);
  const size_t blocks = (size_t)sm_count() * 8;
  return tiles > blocks ? tiles : blocks;
}
std::atomic<int> g_wide_mode{0};   // 0: tensor-core products (tcgen05, 3 x TF32; default), 1: fp32 SIMT products

// k-splits of a weight-gradient product on the persistent tensor-core kernel: tiles * splits work items are walked by one CTA per
// SM, so the time is (waves of items) x (k-blocks per item + the fixed cost of an item, about a dozen k-blocks' worth of
// prologue, drain and epilogue); take the split count that minimises it
static int pick_splits(int tiles, int K) {
  const int sms = sm_count(), kblocks = (K + 31) / 32;
  int best = 1;
  long long best_cost = -1;
  for (int s = 1; s <= kMaxSplits && s <= kblocks; ++s) {
    const long long waves = ((long long)tiles * s + sms - 1) / sms;
    const long long cost = waves * ((kblocks + s - 1) / s + 12);
    if (best_cost < 0 || cost < best_cost) best = s, best_cost = cost;
  }
  return best;
}

static int launch_gemm(const Gemm& g, int splits, cudaStream_t stream) {
  if (g_wide_mode == 0) return launch_gemm_tc(g, splits, stream);
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, splits);
  CVF_LAUNCH(K_AE_STEP, stream, sgemm_kernel<<<grid, 256, 0, stream>>>(g));
  CVF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace wide

int wide_ae_set_mode(int mode) {
  if (mode != 0 && mode != 1) return CVF_E_ARG;
  wide::g_wide_mode = mode;
  return 0;
}

// bytes of workspace the layer-wise path wants for a batch of B frames (it works in chunks of frames that fit)
size_t wide_ae_workspace_bytes(const NetPlan& np, long long B) {
  wide::WidePlan P;
  wide::make_plan(np, &P);
  long long chunk = B < wide::kChunkFrames ? B : wide::kChunkFrames;
  chunk = (chunk + 127) / 128 * 128;
  const size_t per_frame = (P.act_floats_per_frame + P.delta_floats_per_frame + P.img_floats_per_frame) * sizeof(float);
  return 8192 + (P.w_floats + P.wimg_floats) * sizeof(float) + (size_t)wide::kMaxSplits * P.dw_max_floats * sizeof(float) +
         (size_t)chunk * per_frame + wide::part_slots(P) * 2 * sizeof(double) + 4096;
}

int wide_ae_step(const NetPlan& np, const float* feat, const float* w, long long B, const float* params, double* sums_out,
                 double* grad_out, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  using namespace wide;
  WidePlan P;
  make_plan(np, &P);
  const int L = P.L;
  // carve the workspace
  char* base = (char*)workspace;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base + off;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  double* part = (double*)take(part_slots(P) * 2 * sizeof(double));
  float* Wp = (float*)take(P.w_floats * sizeof(float));
  float* Wimg = (float*)take(P.wimg_floats * sizeof(float));
  float* dWp = (float*)take((size_t)kMaxSplits * P.dw_max_floats * sizeof(float));
  if (off >= ws_bytes) {
    set_error("workspace too small for the layer-wise autoencoder path: %zu bytes", ws_bytes);
    return CVF_E_WORKSPACE;
  }
  const size_t per_frame = (P.act_floats_per_frame + P.delta_floats_per_frame + P.img_floats_per_frame) * sizeof(float);
  long long chunk = (long long)((ws_bytes - off - 2048) / per_frame);
  chunk = chunk / 128 * 128;
  if (chunk > kChunkFrames) chunk = kChunkFrames;
  if (chunk < 128) {
    set_error("workspace too small for the layer-wise autoencoder path: %zu bytes leave no room for a 128-frame chunk", ws_bytes);
    return CVF_E_WORKSPACE;
  }
  float* acts[kMaxLayers + 1];
  char* abase = take((size_t)chunk * P.act_floats_per_frame * sizeof(float));
  {
    size_t o = 0;
    for (int l = 0; l <= L; ++l) acts[l] = (float*)abase + o, o += (size_t)chunk * P.ld[l];
  }
  float* dbuf = (float*)take((size_t)chunk * P.delta_floats_per_frame * sizeof(float));
  float* delta[2] = {dbuf, dbuf + (size_t)chunk * (P.delta_floats_per_frame / 2)};
  float* ibuf = (float*)take((size_t)chunk * P.img_floats_per_frame * sizeof(float));
  const bool tc = g_wide_mode == 0;
  // operand images of this chunk (tensor-core path): activations K-major / transposed, deltas as two ping-pong pairs
  float *imgK[kMaxLayers], *imgT[kMaxLayers], *dK[2], *dT[2];
  {
    size_t o = 0;
    for (int l = 0; l < L; ++l) imgK[l] = ibuf + (size_t)chunk * P.ak_off[l], imgT[l] = ibuf + (size_t)chunk * P.at_off[l];
    o = (size_t)chunk * (P.img_floats_per_frame - 2 * (P.dk_floats + P.dt_floats));
    for (int b = 0; b < 2; ++b) {
      dK[b] = ibuf + o, o += (size_t)chunk * P.dk_floats;
      dT[b] = ibuf + o, o += (size_t)chunk * P.dt_floats;
    }
  }
  // the tensor-core path reads the caller's features in place when their rows are 16-byte aligned; the SIMT products want the
  // padded copy with the ones column
  const bool in_place = tc && P.dims[0] % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0;

  for (int l = 0; l < L; ++l) {
    const int n = P.dims[l + 1] * P.ld[l];
    CVF_LAUNCH(K_AE_STEP, stream,
               pad_weights_kernel<<<(n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256, 256, 0, stream>>>(params + P.gw_off[l], P.dims[l + 1],
                                                                                                     P.dims[l], Wp + P.w_off[l], P.ld[l]));
  }
  if (tc) {   // weight images: B operand of the forward products and (layers >= 1, training only) of the back-propagation products
    for (int l = 0; l < L; ++l) {
      int e = launch_tile_image(Wp + P.w_off[l], P.ld[l], 1, P.dims[l + 1], P.dims[l], 0, Wimg + P.wf_off[l], stream);
      if (e) return e;
      if (grad_out && l >= 1) {
        e = launch_tile_image(Wp + P.w_off[l], P.ld[l], 0, P.dims[l], P.dims[l + 1], 0, Wimg + P.wb_off[l], stream);
        if (e) return e;
      }
    }
  }
  if (grad_out) CVF_LAUNCH(K_AE_STEP, stream, zero_doubles_kernel<<<64, 256, 0, stream>>>(grad_out, np.n_params));
  CVF_CUDA(cudaGetLastError());

  bool first = true;
  for (long long f0 = 0; f0 < B; f0 += chunk) {
    const int M = (int)(B - f0 < chunk ? B - f0 : chunk);
    const int fblocks = (M + 31) / 32;   // k-blocks of an operand whose K index is the frame
    const float* in0 = in_place ? feat + (size_t)f0 * P.dims[0] : acts[0];
    const int ld0 = in_place ? P.dims[0] : P.ld[0];
    if (!in_place) {
      const long long n = (long long)M * P.ld[0];
      const int grid = (int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
      CVF_LAUNCH(K_AE_STEP, stream, stage_input_kernel<<<grid, 256, 0, stream>>>(feat, f0, M, P.dims[0], acts[0], P.ld[0]));
    }
    if (tc) {   // the input as the A operand of the first product and (training) as the B operand of dW_1, with its ones row
      int e = launch_tile_image(in0, ld0, 1, M, P.dims[0], 0, imgK[0], stream);
      if (e) return e;
      if (grad_out) {
        e = launch_tile_image(in0, ld0, 0, P.dims[0], M, P.dims[0], imgT[0], stream);
        if (e) return e;
      }
    } else {
      for (int l = 1; l < L; ++l) {   // ones column / zero padding of the hidden activation buffers
        const long long n = (long long)M * (P.ld[l] - P.dims[l]);
        const int grid = (int)((n + 255) / 256 > 1024 ? 1024 : (n + 255) / 256);
        CVF_LAUNCH(K_AE_STEP, stream, set_pad_kernel<<<grid, 256, 0, stream>>>(acts[l], M, P.dims[l], P.ld[l]));
      }
    }
    // the tensor-core path folds the loss into the last product's epilogue: delta_L leaves it as images, A_L is never stored
    const bool fused_loss = tc && !P.act[L - 1];
    int loss_slots = 0;
    for (int l = 0; l < L; ++l) {   // forward
      Gemm g;
      memset(&g, 0, sizeof(g));
      g.A = l == 0 ? in0 : acts[l], g.lda = l == 0 ? ld0 : P.ld[l], g.a_kcontig = 1;
      g.B = Wp + P.w_off[l], g.ldb = P.ld[l], g.b_kcontig = 1;
      g.C = acts[l + 1], g.ldc = P.ld[l + 1];
      g.M = M, g.N = P.dims[l + 1], g.K = P.dims[l], g.k_per_split = g.K;
      g.epi = P.act[l] ? EPI_BIAS_TANH : EPI_BIAS;
      g.bias = params + P.gb_off[l];
      if (tc) {
        g.a_img = imgK[l], g.a_img_kblocks = (P.dims[l] + 31) / 32;
        g.b_img = Wimg + P.wf_off[l], g.b_img_kblocks = (P.dims[l] + 31) / 32;
        if (l + 1 < L) {   // the next layer and the backward pass read A_{l+1} from its images: no row-major copy
          g.C = nullptr;
          g.c_img_k = imgK[l + 1], g.c_img_k_kblocks = (P.dims[l + 1] + 31) / 32;
          if (grad_out) g.c_img_t = imgT[l + 1], g.c_img_t_kblocks = fblocks, g.c_img_t_ones = P.dims[l + 1];
        } else if (fused_loss) {
          g.epi = EPI_BIAS_LOSS;
          g.loss_in = in0, g.loss_ld = ld0, g.loss_w = w + f0, g.loss_part = part;
          g.C = nullptr;
          loss_slots = gemm_tc_tiles(g, 1);
          if (grad_out) {
            g.c_img_t = dT[0], g.c_img_t_kblocks = fblocks;
            if (L >= 2) g.c_img_k = dK[0], g.c_img_k_kblocks = (P.dims[L] + 31) / 32;
          }
        }
      }
      int e = launch_gemm(g, 1, stream);
      if (e) return e;
    }
    if (!fused_loss) {
      loss_slots = sm_count() * 8;
      if ((M + 7) / 8 < loss_slots) loss_slots = (M + 7) / 8;
      CVF_LAUNCH(K_AE_STEP, stream,
                 loss_delta_kernel<<<loss_slots, 256, 0, stream>>>(acts[L], in0, w, f0, M, P.dims[L], P.ld[L], ld0, part));
    }
    CVF_LAUNCH(K_REDUCE, stream, add_sums_kernel<<<1, 256, 0, stream>>>(part, loss_slots, sums_out, first ? 1 : 0));
    CVF_CUDA(cudaGetLastError());
    first = false;
    if (!grad_out) continue;
    // backward: delta_L lives in acts[L]; lower deltas ping-pong in dbuf (SIMT path) / in the image pairs (tensor-core path)
    const float* dcur = acts[L];
    int cur_ld = P.ld[L], cur = 0;
    if (tc && !fused_loss) {
      int e = launch_tile_image(dcur, cur_ld, 0, P.dims[L], M, 0, dT[0], stream);
      if (e) return e;
      if (L >= 2) {
        e = launch_tile_image(dcur, cur_ld, 1, M, P.dims[L], 0, dK[0], stream);
        if (e) return e;
      }
    }
    for (int l = L - 1; l >= 0; --l) {
      // [dW | db] = delta^T [A_l | 1], split over the frames
      {
        const int rows = P.dims[l + 1], cols = P.dims[l] + 1;
        const int tiles = ((rows + BM - 1) / BM) * ((cols + BN - 1) / BN);
        int splits = tc ? pick_splits(tiles, M) : (2 * sm_count() + tiles - 1) / tiles;
        if (splits > kMaxSplits) splits = kMaxSplits;
        int kps = ((M + splits - 1) / splits + 31) / 32 * 32;
        splits = (M + kps - 1) / kps;
        Gemm g;
        memset(&g, 0, sizeof(g));
        g.A = dcur, g.lda = cur_ld, g.a_kcontig = 0;          // Aop[i][f] = delta[f][i]
        g.B = acts[l], g.ldb = P.ld[l], g.b_kcontig = 0;      // Bop[f][j] = A_l[f][j]
        if (tc) g.a_img = dT[cur], g.b_img = imgT[l], g.a_img_kblocks = g.b_img_kblocks = fblocks;
        g.C = dWp, g.ldc = P.ld[l], g.c_split_stride = (long long)rows * P.ld[l];
        g.M = rows, g.N = cols, g.K = M, g.k_per_split = kps;
        g.epi = EPI_NONE;
        int e = launch_gemm(g, splits, stream);
        if (e) return e;
        const int n = rows * cols;
        CVF_LAUNCH(K_REDUCE, stream,
                   accumulate_grad_kernel<<<(n + 255) / 256 > 2048 ? 2048 : (n + 255) / 256, 256, 0, stream>>>(
                       dWp, splits, (long long)rows * P.ld[l], rows, P.dims[l], P.ld[l], grad_out, P.gw_off[l], P.gb_off[l]));
        CVF_CUDA(cudaGetLastError());
      }
      if (l == 0) break;
      // delta_{l} (for layer l's output, i.e. A_l) = (delta_{l+1} W_{l+1}) .* (1 - A_l^2) when layer l has an activation
      {
        float* dnext = delta[l & 1];
        Gemm g;
        memset(&g, 0, sizeof(g));
        g.A = dcur, g.lda = cur_ld, g.a_kcontig = 1;                    // Aop[f][i] = delta[f][i]
        g.B = Wp + P.w_off[l], g.ldb = P.ld[l], g.b_kcontig = 0;        // Bop[i][j] = W[i][j]
        g.C = dnext, g.ldc = P.ld[l];
        g.M = M, g.N = P.dims[l], g.K = P.dims[l + 1], g.k_per_split = g.K;
        g.epi = P.act[l - 1] ? EPI_MUL_OM : EPI_NONE;
        g.act = acts[l];
        if (tc) {   // the new delta leaves the product as images only: nothing reads it as a row-major matrix
          g.a_img = dK[cur], g.a_img_kblocks = (P.dims[l + 1] + 31) / 32;
          g.b_img = Wimg + P.wb_off[l], g.b_img_kblocks = (P.dims[l + 1] + 31) / 32;
          g.C = nullptr;
          g.act_img = imgK[l], g.act_img_kblocks = (P.dims[l] + 31) / 32;
          g.c_img_t = dT[cur ^ 1], g.c_img_t_kblocks = fblocks;
          if (l >= 2) g.c_img_k = dK[cur ^ 1], g.c_img_k_kblocks = (P.dims[l] + 31) / 32;
        }
        int e = launch_gemm(g, 1, stream);
        if (e) return e;
        dcur = dnext, cur_ld = P.ld[l], cur ^= 1;
      }
    }
  }
  return 0;
}

}  // namespace cvf
