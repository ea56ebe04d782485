// cvf_ae_fast.cu -- AutoEncoderTask.weighted_MSE_loss + backward (reference core.py:652-666, 708) for the notebook-sized
// autoencoder  encoder [d, 20, 20, 20, e] / decoder [e, 10, 10, d], e = 1, 2, 3  (examples/dipeptide/main.ipynb:434), organised like the
// eigenfunction fast path (cvf_eigen_fast.cu) around what the SM sustains:
//
//   prep   : the caller's [B][d] features -> frame-minor rows R[0..d) (row = feature, column = frame)
//   main   : forward through the seven layers and the reverse sweep of the deltas, THREAD-PRIVATELY (two CTAs per SM, two frames per
//            thread so that every weight fetched from shared memory feeds two FFMA2, activations in registers, no barrier
//            between layers); the activations a_1..a_6 and the deltas delta_1..delta_7 go to frame-minor rows of R, the
//            weighted squared error to fp64 partial sums
//   dw     : every weight / bias gradient dW_l = sum_f delta_l (x) a_{l-1}, db_l = sum_f delta_l as products over the frames
//            with the accumulators in registers (4 x 12 per lane); the seven layers are packed into "types" of at most 32
//            lanes, a pair of warps shares two cp.async staging buffers per type and run of tiles
// Sums have one owner each (per-warp fp64 blocks, folded in a fixed order): the result is deterministic.
#include <atomic>
#include <string.h>

#include "cvf_common.cuh"
#include "cvf_tma.cuh"

namespace cvf {
namespace aefast {

typedef unsigned long long u64;
constexpr int kRP = 36;        // floats per 32-frame operand row in shared memory (16-byte aligned, bank skew 4 per row)
constexpr int kRun = 16;       // tiles of 32 frames per work item of the dw kernel
constexpr int kTile = 256;     // frames per tile of the main kernel: two CTAs per SM, so that one computes while the other
constexpr int kThreads = 128;  // waits for its (single-buffered) tile
constexpr int kMaxTypes = 4;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)), "l"(*reinterpret_cast<u64*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 dup(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Plan {
  int L, d, drp;
  int dims[kMaxLayers + 1];
  int aoff[kMaxLayers + 1];   // first row in R of a_l (a_0 = the input), l = 0 .. L-1
  int doff[kMaxLayers + 1];   // first row in R of delta_l, l = 1 .. L
  int n_rows;
  int gw_off[kMaxLayers], gb_off[kMaxLayers], n_params;
  // dw kernel: layers packed into types of at most 32 lanes; per (type, slot): layer, first lane, buffer rows of its operands
  int n_types, type_n[kMaxTypes], type_layer[kMaxTypes][kMaxLayers], type_lane0[kMaxTypes][kMaxLayers];
  int type_xrow[kMaxTypes][kMaxLayers], type_zrow[kMaxTypes][kMaxLayers], rows_buf;
  int img_floats;
  long long B, Bp;
  float* R;          // [n_rows][Bp]
  float* img;        // shared-memory image of the parameters
  double* part_loss; // [grid of main][2]
  double* part_dw;   // [warps of dw][n_params]
};

// ---- shared-memory image of the parameters (floats).  k-major blocks WlT[in][outp] feed the forward layers, natural blocks
// Wln[out][inp] the reverse sweep; every row is padded to a multiple of 4 floats.
template <int E, int EC, int G>
struct Img {
  static constexpr int e4 = (EC + 3) & ~3, G12 = (G + 3) & ~3;
  __host__ __device__ static int w1t() { return 0; }
  __host__ __device__ static int b1(int d) { return d * E; }
  __host__ __device__ static int w2t(int d) { return b1(d) + E; }
  __host__ __device__ static int b2(int d) { return w2t(d) + E * E; }
  __host__ __device__ static int w3t(int d) { return b2(d) + E; }
  __host__ __device__ static int b3(int d) { return w3t(d) + E * E; }
  __host__ __device__ static int w4t(int d) { return b3(d) + E; }            // [E][e4]
  __host__ __device__ static int b4(int d) { return w4t(d) + E * e4; }
  __host__ __device__ static int w5t(int d) { return b4(d) + e4; }           // [EC][G12]
  __host__ __device__ static int b5(int d) { return w5t(d) + EC * G12; }
  __host__ __device__ static int w6t(int d) { return b5(d) + G12; }          // [G][G12]
  __host__ __device__ static int b6(int d) { return w6t(d) + G * G12; }
  __host__ __device__ static int w7t(int d) { return b6(d) + G12; }          // [G][drp]
  __host__ __device__ static int b7(int d, int drp) { return w7t(d) + G * drp; }
  __host__ __device__ static int w7n(int d, int drp) { return b7(d, drp) + drp; }    // [d][G12]
  __host__ __device__ static int w6n(int d, int drp) { return w7n(d, drp) + d * G12; }   // [G][G12]
  __host__ __device__ static int w5n(int d, int drp) { return w6n(d, drp) + G * G12; }   // [G][e4]
  __host__ __device__ static int w4n(int d, int drp) { return w5n(d, drp) + G * e4; }    // [EC][E]
  __host__ __device__ static int w3n(int d, int drp) { return w4n(d, drp) + EC * E; }    // [E][E]
  __host__ __device__ static int w2n(int d, int drp) { return w3n(d, drp) + E * E; }     // [E][E]
  __host__ __device__ static int floats(int d, int drp) { return w2n(d, drp) + E * E; }
};

// k-major block: dst[i * outp + o] = W[o][i]; natural block: dst[o * inp + i] = W[o][i]; padding zero
__device__ void pack_kmajor(float* dst, const float* W, int n_out, int n_in, int outp, int tid, int nt) {
  for (int t = tid; t < n_in * outp; t += nt) {
    const int i = t / outp, o = t - i * outp;
    dst[t] = o < n_out ? W[o * n_in + i] : 0.0f;
  }
}
__device__ void pack_natural(float* dst, const float* W, int n_out, int n_in, int inp, int tid, int nt) {
  for (int t = tid; t < n_out * inp; t += nt) {
    const int o = t / inp, i = t - o * inp;
    dst[t] = i < n_in ? W[o * n_in + i] : 0.0f;
  }
}
__device__ void pack_vec(float* dst, const float* b, int n, int np, int tid, int nt) {
  for (int t = tid; t < np; t += nt) dst[t] = t < n ? b[t] : 0.0f;
}

template <int E, int EC, int G>
__global__ void pack_kernel(const Plan P, const float* __restrict__ params) {
  typedef Img<E, EC, G> I;
  const int d = P.d, drp = P.drp, tid = threadIdx.x, nt = blockDim.x;
  float* g = P.img;
  const float* W[7];
  const float* b[7];
  for (int l = 0; l < 7; ++l) W[l] = params + P.gw_off[l], b[l] = params + P.gb_off[l];
  switch (blockIdx.x) {
    case 0: pack_kmajor(g + I::w1t(), W[0], E, d, E, tid, nt); pack_vec(g + I::b1(d), b[0], E, E, tid, nt); break;
    case 1: pack_kmajor(g + I::w2t(d), W[1], E, E, E, tid, nt); pack_vec(g + I::b2(d), b[1], E, E, tid, nt);
            pack_natural(g + I::w2n(d, drp), W[1], E, E, E, tid, nt); break;
    case 2: pack_kmajor(g + I::w3t(d), W[2], E, E, E, tid, nt); pack_vec(g + I::b3(d), b[2], E, E, tid, nt);
            pack_natural(g + I::w3n(d, drp), W[2], E, E, E, tid, nt); break;
    case 3: pack_kmajor(g + I::w4t(d), W[3], EC, E, I::e4, tid, nt); pack_vec(g + I::b4(d), b[3], EC, I::e4, tid, nt);
            pack_natural(g + I::w4n(d, drp), W[3], EC, E, E, tid, nt); break;
    case 4: pack_kmajor(g + I::w5t(d), W[4], G, EC, I::G12, tid, nt); pack_vec(g + I::b5(d), b[4], G, I::G12, tid, nt);
            pack_natural(g + I::w5n(d, drp), W[4], G, EC, I::e4, tid, nt); break;
    case 5: pack_kmajor(g + I::w6t(d), W[5], G, G, I::G12, tid, nt); pack_vec(g + I::b6(d), b[5], G, I::G12, tid, nt);
            pack_natural(g + I::w6n(d, drp), W[5], G, G, I::G12, tid, nt); break;
    default: pack_kmajor(g + I::w7t(d), W[6], d, G, drp, tid, nt); pack_vec(g + I::b7(d, drp), b[6], d, drp, tid, nt);
             pack_natural(g + I::w7n(d, drp), W[6], d, G, I::G12, tid, nt); break;
  }
}

// ------------------------------------------------------------------------------------------------ prep
// [B][d] -> R[0..d)[Bp]: 128 frames per CTA pass through a shared-memory tile of odd row stride (coalesced both ways);
// padding frames repeat the last frame (their weight is 0)
__global__ void __launch_bounds__(128) prep_kernel(const Plan P, const float* __restrict__ feat) {
  extern __shared__ __align__(16) float st[];
  const int tid = threadIdx.x, d = P.d, S = d | 1;
  const long long n_tiles = P.Bp / 128;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long f0 = tile * 128;
    if (f0 + 128 <= P.B) {
      // contiguous tile: element i = tid + 128 m belongs to frame i / d; (frame, feature) advance incrementally
      const float* src = feat + (size_t)f0 * d;
      const int df = 128 / d, dj = 128 - df * d;
      int f = tid / d, j = tid - f * d;
      for (int i = tid; i < 128 * d; i += 128) {
        st[f * S + j] = __ldg(src + i);
        f += df, j += dj;
        if (j >= d) j -= d, ++f;
      }
    } else {
      for (int i = tid; i < 128 * d; i += 128) {
        const int f = i / d, j = i - f * d;
        const long long fr = min(f0 + f, P.B - 1);
        st[f * S + j] = __ldg(feat + (size_t)fr * d + j);
      }
    }
    __syncthreads();
    for (int r = 0; r < d; ++r) P.R[(size_t)r * P.Bp + f0 + tid] = st[tid * S + r];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ main
// z[f][..] += WT[kk][..] * x[f][kk] for both frames: one 128-bit weight load feeds four FFMA2
template <int IN, int OUTP>
__device__ __forceinline__ void dense2(float2 (&z)[2][OUTP / 2], const float* __restrict__ WT, const float (&x)[2][IN]) {
#pragma unroll
  for (int kk = 0; kk < IN; ++kk) {
    const float2 x0 = dup(x[0][kk]), x1 = dup(x[1][kk]);
#pragma unroll
    for (int q = 0; q < OUTP / 4; ++q) {
      const float4 wv = ld4(WT + kk * OUTP + 4 * q);
      z[0][2 * q] = ffma2(make_float2(wv.x, wv.y), x0, z[0][2 * q]);
      z[0][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), x0, z[0][2 * q + 1]);
      z[1][2 * q] = ffma2(make_float2(wv.x, wv.y), x1, z[1][2 * q]);
      z[1][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), x1, z[1][2 * q + 1]);
    }
  }
}
template <int OUTP>
__device__ __forceinline__ void load_bias2(float2 (&z)[2][OUTP / 2], const float* __restrict__ b) {
#pragma unroll
  for (int j = 0; j < OUTP / 2; ++j) z[0][j] = z[1][j] = lds2(b + 2 * j);
}
// rows [row0, row0 + N) of R at this thread's two frames
template <int N>
__device__ __forceinline__ void store_rows2(float* __restrict__ R, long long Bp, int row0, long long col, const float (&v)[2][N]) {
#pragma unroll
  for (int j = 0; j < N; ++j) *reinterpret_cast<float2*>(R + (size_t)(row0 + j) * Bp + col) = make_float2(v[0][j], v[1][j]);
}
template <int N>
__device__ __forceinline__ void load_rows2(const float* __restrict__ R, long long Bp, int row0, long long col, float (&v)[2][N]) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float2 t = __ldcg(reinterpret_cast<const float2*>(R + (size_t)(row0 + j) * Bp + col));
    v[0][j] = t.x, v[1][j] = t.y;
  }
}
template <int N, int NP>
__device__ __forceinline__ void tanh2(float (&a)[2][N], const float2 (&z)[2][NP / 2]) {
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      if (j + 1 < N) {   // two at a time, packed (bit-identical to cvf_tanh)
        const float2 t = cvf_tanh2(z[f][j >> 1]);
        a[f][j] = t.x, a[f][j + 1] = t.y;
      } else {
        a[f][j] = cvf_tanh(z[f][j >> 1].x);
      }
    }
}
template <int N, int NP>
__device__ __forceinline__ void unpack2(float (&a)[2][N], const float2 (&z)[2][NP / 2]) {
#pragma unroll
  for (int f = 0; f < 2; ++f)
#pragma unroll
    for (int j = 0; j < N; ++j) a[f][j] = (j & 1) ? z[f][j >> 1].y : z[f][j >> 1].x;
}

template <int E, int EC, int G>
__global__ void __launch_bounds__(kThreads, 2) main_kernel(const Plan P, const float* __restrict__ w, int grad) {
  extern __shared__ __align__(16) float sm[];
  __shared__ double red[2][kThreads / 32];
  typedef Img<E, EC, G> I;
  constexpr int e4 = I::e4, G12 = I::G12, F = kTile;
  const int tid = threadIdx.x, d = P.d, drp = P.drp;
  float* W = sm;
  float* tile = sm + P.img_floats;   // [d][F]
  for (int i = tid; i < P.img_floats; i += kThreads) W[i] = P.img[i];
  const long long n_tiles = P.Bp / F;
  const int c0 = 2 * tid;
  double loss_acc = 0.0, w_acc = 0.0;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long f0 = t * F, col = f0 + c0;
    __syncthreads();
    for (int i = tid; i < d * (F / 4); i += kThreads) {
      const int r = i / (F / 4), c4 = i - r * (F / 4);
      st4(tile + r * F + 4 * c4, __ldg(reinterpret_cast<const float4*>(P.R + (size_t)r * P.Bp + f0) + c4));
    }
    __syncthreads();
    {
      // pull the next tile into L2 while this one is worked on
      const long long tn = t + gridDim.x;
      if (tn < n_tiles)
        for (int i = tid; i < d * (F / 32); i += kThreads) {
          const int r = i / (F / 32), c = i - r * (F / 32);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(P.R + (size_t)r * P.Bp + tn * F + 32 * c));
        }
    }
    float wf[2];
    wf[0] = col < P.B ? __ldg(w + col) : 0.0f;
    wf[1] = col + 1 < P.B ? __ldg(w + col + 1) : 0.0f;
    w_acc += (double)wf[0] + (double)wf[1];
    // ---- encoder (nn.py:52-57): three tanh layers E wide, a linear layer to the EC-dimensional code
    float a[2][E];
    {
      float2 z[2][E / 2];
      load_bias2<E>(z, W + I::b1(d));
#pragma unroll 2
      for (int kk = 0; kk < d; ++kk) {
        const float2 x = lds2(tile + kk * F + c0);
        const float2 x0 = dup(x.x), x1 = dup(x.y);
        const float* wr = W + I::w1t() + kk * E;
#pragma unroll
        for (int q = 0; q < E / 4; ++q) {
          const float4 wv = ld4(wr + 4 * q);
          z[0][2 * q] = ffma2(make_float2(wv.x, wv.y), x0, z[0][2 * q]);
          z[0][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), x0, z[0][2 * q + 1]);
          z[1][2 * q] = ffma2(make_float2(wv.x, wv.y), x1, z[1][2 * q]);
          z[1][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), x1, z[1][2 * q + 1]);
        }
      }
      tanh2<E, E>(a, z);
      if (grad) store_rows2<E>(P.R, P.Bp, P.aoff[1], col, a);
    }
#pragma unroll 1
    for (int l = 2; l <= 3; ++l) {
      float2 z[2][E / 2];
      load_bias2<E>(z, W + (l == 2 ? I::b2(d) : I::b3(d)));
      dense2<E, E>(z, W + (l == 2 ? I::w2t(d) : I::w3t(d)), a);
      tanh2<E, E>(a, z);
      if (grad) store_rows2<E>(P.R, P.Bp, P.aoff[l], col, a);
    }
    float a4[2][EC];
    {
      float2 z[2][e4 / 2];
      load_bias2<e4>(z, W + I::b4(d));
      dense2<E, e4>(z, W + I::w4t(d), a);
      unpack2<EC, e4>(a4, z);
      if (grad) store_rows2<EC>(P.R, P.Bp, P.aoff[4], col, a4);
    }
    // ---- decoder: two tanh layers G wide, a linear layer back to d
    float a5[2][G], a6[2][G];
    {
      float2 z[2][G12 / 2];
      load_bias2<G12>(z, W + I::b5(d));
      dense2<EC, G12>(z, W + I::w5t(d), a4);
      tanh2<G, G12>(a5, z);
      if (grad) store_rows2<G>(P.R, P.Bp, P.aoff[5], col, a5);
      load_bias2<G12>(z, W + I::b6(d));
      dense2<G, G12>(z, W + I::w6t(d), a5);
      tanh2<G, G12>(a6, z);
      if (grad) store_rows2<G>(P.R, P.Bp, P.aoff[6], col, a6);
    }
    // ---- output in chunks of 12 coordinates: error, loss, delta_7 = 2 w (out - x) (core.py:666), h_6 = W_7^T delta_7
    float2 h6[2][G12 / 2];
#pragma unroll
    for (int j = 0; j < G12 / 2; ++j) h6[0][j] = h6[1][j] = make_float2(0.f, 0.f);
    float lacc[2] = {0.f, 0.f};
    for (int c = 0; c < drp / 12; ++c) {
      float2 acc[2][6];
#pragma unroll
      for (int p = 0; p < 6; ++p) acc[0][p] = acc[1][p] = lds2(W + I::b7(d, drp) + 12 * c + 2 * p);
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const float2 x0 = dup(a6[0][i]), x1 = dup(a6[1][i]);
        const float* wr = W + I::w7t(d) + i * drp + 12 * c;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const float4 wv = ld4(wr + 4 * q);
          acc[0][2 * q] = ffma2(make_float2(wv.x, wv.y), x0, acc[0][2 * q]);
          acc[0][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), x0, acc[0][2 * q + 1]);
          acc[1][2 * q] = ffma2(make_float2(wv.x, wv.y), x1, acc[1][2 * q]);
          acc[1][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), x1, acc[1][2 * q + 1]);
        }
      }
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const int r = 12 * c + j;
        if (r < d) {
          const float2 x = lds2(tile + r * F + c0);
          const float o0 = (j & 1) ? acc[0][j >> 1].y : acc[0][j >> 1].x, o1 = (j & 1) ? acc[1][j >> 1].y : acc[1][j >> 1].x;
          const float e0 = o0 - x.x, e1 = o1 - x.y;
          lacc[0] = fmaf(e0, e0, lacc[0]), lacc[1] = fmaf(e1, e1, lacc[1]);
          if (grad) {
            const float d0 = 2.0f * wf[0] * e0, d1 = 2.0f * wf[1] * e1;
            *reinterpret_cast<float2*>(P.R + (size_t)(P.doff[7] + r) * P.Bp + col) = make_float2(d0, d1);
            const float2 g0 = dup(d0), g1 = dup(d1);
            const float* wn = W + I::w7n(d, drp) + r * G12;
#pragma unroll
            for (int q = 0; q < G12 / 4; ++q) {
              const float4 wv = ld4(wn + 4 * q);
              h6[0][2 * q] = ffma2(make_float2(wv.x, wv.y), g0, h6[0][2 * q]);
              h6[0][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), g0, h6[0][2 * q + 1]);
              h6[1][2 * q] = ffma2(make_float2(wv.x, wv.y), g1, h6[1][2 * q]);
              h6[1][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), g1, h6[1][2 * q + 1]);
            }
          }
        }
      }
    }
    loss_acc += (double)wf[0] * (double)lacc[0] + (double)wf[1] * (double)lacc[1];
    if (!grad) continue;
    // ---- reverse sweep: delta_l = (W_{l+1}^T delta_{l+1}) .* (1 - a_l^2) for the tanh layers, delta_4 = W_5^T delta_5
    float dl[2][G];
    {
      float h[2][G];
      unpack2<G, G12>(h, h6);
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int j = 0; j < G; ++j) dl[f][j] = h[f][j] * fmaf(-a6[f][j], a6[f][j], 1.0f);
      store_rows2<G>(P.R, P.Bp, P.doff[6], col, dl);
      float2 z[2][G12 / 2];
#pragma unroll
      for (int j = 0; j < G12 / 2; ++j) z[0][j] = z[1][j] = make_float2(0.f, 0.f);
      dense2<G, G12>(z, W + I::w6n(d, drp), dl);      // rows o of W_6 [G][G12]: h_5[i] += W_6[o][i] delta_6[o]
      unpack2<G, G12>(h, z);
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int j = 0; j < G; ++j) dl[f][j] = h[f][j] * fmaf(-a5[f][j], a5[f][j], 1.0f);
      store_rows2<G>(P.R, P.Bp, P.doff[5], col, dl);
    }
    float d4[2][EC];
    {
      float2 z[2][e4 / 2];
#pragma unroll
      for (int j = 0; j < e4 / 2; ++j) z[0][j] = z[1][j] = make_float2(0.f, 0.f);
      dense2<G, e4>(z, W + I::w5n(d, drp), dl);       // W_5 [G][e4]
      unpack2<EC, e4>(d4, z);
      store_rows2<EC>(P.R, P.Bp, P.doff[4], col, d4);
    }
    float de[2][E];
    {
      float2 z[2][E / 2];
#pragma unroll
      for (int j = 0; j < E / 2; ++j) z[0][j] = z[1][j] = make_float2(0.f, 0.f);
      dense2<EC, E>(z, W + I::w4n(d, drp), d4);       // W_4 [EC][E]
      float h[2][E];
      unpack2<E, E>(h, z);
      load_rows2<E>(P.R, P.Bp, P.aoff[3], col, a);
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int j = 0; j < E; ++j) de[f][j] = h[f][j] * fmaf(-a[f][j], a[f][j], 1.0f);
      store_rows2<E>(P.R, P.Bp, P.doff[3], col, de);
    }
#pragma unroll 1
    for (int l = 2; l >= 1; --l) {
      float2 z[2][E / 2];
#pragma unroll
      for (int j = 0; j < E / 2; ++j) z[0][j] = z[1][j] = make_float2(0.f, 0.f);
      dense2<E, E>(z, W + (l == 2 ? I::w3n(d, drp) : I::w2n(d, drp)), de);   // h_l = W_{l+1}^T delta_{l+1}
      float h[2][E];
      unpack2<E, E>(h, z);
      load_rows2<E>(P.R, P.Bp, P.aoff[l], col, a);
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int j = 0; j < E; ++j) de[f][j] = h[f][j] * fmaf(-a[f][j], a[f][j], 1.0f);
      store_rows2<E>(P.R, P.Bp, P.doff[l], col, de);
    }
  }
  // fp64 partial sums of the CTA: sum w |e|^2, sum w
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    w_acc += __shfl_xor_sync(0xffffffffu, w_acc, o);
  }
  if ((tid & 31) == 0) red[0][tid >> 5] = loss_acc, red[1][tid >> 5] = w_acc;
  __syncthreads();
  if (tid < 2) {
    double s = 0.0;
    for (int q = 0; q < kThreads / 32; ++q) s += red[tid][q];
    P.part_loss[(size_t)blockIdx.x * 2 + tid] = s;
  }
}

// ------------------------------------------------------------------------------------------------ dw
// acc[j][i] += sum_f X[j][f] Z[i][f] over frames [fbeg, fend) of the staged rows; bsum[j] += sum_f X[j][f] when `bias`
__device__ __forceinline__ void outer_tile_b(float2 (&acc)[4][12], float (&bsum)[4], const float* __restrict__ X,
                                             const float* __restrict__ Z, int sx, int sz, int fbeg, int fend, bool bias) {
#pragma unroll 1
  for (int f = fbeg; f < fend; f += 4) {
    float4 x[4], z[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = ld4(X + j * sx + f);
#pragma unroll
    for (int i = 0; i < 12; ++i) z[i] = ld4(Z + i * sz + f);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 12; ++i) acc[j][i] = ffma2(make_float2(x[j].x, x[j].y), make_float2(z[i].x, z[i].y), acc[j][i]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 12; ++i) acc[j][i] = ffma2(make_float2(x[j].z, x[j].w), make_float2(z[i].z, z[i].w), acc[j][i]);
    if (bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) bsum[j] += (x[j].x + x[j].y) + (x[j].z + x[j].w);
    }
  }
}

__global__ void __launch_bounds__(256, 1) dw_kernel(const Plan P) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
  const int pair = warp >> 1, h = warp & 1, npairs = nw >> 1;
  const int rows_buf = P.rows_buf;
  float* buf0 = sm + (size_t)pair * 2 * rows_buf * kRP;
  for (int i = tid; i < npairs * 2 * rows_buf * kRP; i += nt) sm[i] = 0.0f;
  double* part = P.part_dw + ((size_t)blockIdx.x * nw + warp) * (size_t)P.n_params;
  for (int i = lane; i < P.n_params; i += 32) part[i] = 0.0;
  __syncthreads();
  const long long n_tiles = P.Bp / 32;
  const long long n_runs = (n_tiles + kRun - 1) / kRun, n_items = n_runs * P.n_types;
  const long long stride = (long long)gridDim.x * npairs;
  const int c4 = 4 * (lane & 7);
  // stage the operand rows of every layer of type `ty` for tile t: this warp takes every other row quad
  auto stage = [&](int b, int ty, long long t) {
    float* Rb = buf0 + (size_t)b * rows_buf * kRP;
    for (int s = 0; s < P.type_n[ty]; ++s) {
      const int l = P.type_layer[ty][s];   // layer l maps a_{l-1} (dims[l-1] rows) to delta_l (dims[l] rows)
      const float* srcx = P.R + (size_t)P.doff[l] * P.Bp + t * 32;
      const float* srcz = P.R + (size_t)P.aoff[l - 1] * P.Bp + t * 32;
      float* dx = Rb + P.type_xrow[ty][s] * kRP;
      float* dz = Rb + P.type_zrow[ty][s] * kRP;
      for (int r = (lane >> 3) + 4 * h; r < P.dims[l]; r += 8) cp_async16(dx + r * kRP + c4, srcx + (size_t)r * P.Bp + c4);
      for (int r = (lane >> 3) + 4 * h; r < P.dims[l - 1]; r += 8) cp_async16(dz + r * kRP + c4, srcz + (size_t)r * P.Bp + c4);
    }
    cp_async_commit();
  };
  long long q = (long long)blockIdx.x * npairs + pair;
  int cur = 0;
  if (q < n_items) stage(0, (int)(q % P.n_types), (q / P.n_types) * kRun);
  for (; q < n_items; q += stride) {
    const int ty = (int)(q % P.n_types);
    const long long t0 = (q / P.n_types) * kRun, t1 = t0 + kRun < n_tiles ? t0 + kRun : n_tiles;
    // this lane's block of this type: layer, row / column group
    int l = 0, og = 0, ig = 0, nog = 1, nig = 1, xrow = 0, zrow = 0;
    bool active = false;
    for (int s = 0; s < P.type_n[ty]; ++s) {
      const int ls = P.type_layer[ty][s];
      const int no = (P.dims[ls] + 3) / 4, ni = (P.dims[ls - 1] + 11) / 12;
      const int rel = lane - P.type_lane0[ty][s];
      if (rel >= 0 && rel < no * ni) {
        active = true, l = ls, nog = no, nig = ni, og = rel / ni, ig = rel - og * ni;
        xrow = P.type_xrow[ty][s], zrow = P.type_zrow[ty][s];
      }
    }
    float2 acc[4][12];
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 12; ++i) acc[j][i] = make_float2(0.f, 0.f);
    for (long long t = t0; t < t1; ++t) {
      long long tn = t + 1, qn = q;
      if (tn >= t1) qn = q + stride, tn = (qn / P.n_types) * kRun;
      named_barrier(1 + pair, 64);   // both warps have finished reading the buffer that is staged next
      if (qn < n_items) {
        stage(cur ^ 1, (int)(qn % P.n_types), tn);
        cp_async_wait_group<1>();
      } else {
        cp_async_wait_all();
      }
      named_barrier(1 + pair, 64);   // both halves of the current tile have landed
      const float* Rb = buf0 + (size_t)cur * rows_buf * kRP;
      if (active)
        outer_tile_b(acc, bsum, Rb + (xrow + og) * kRP, Rb + (zrow + ig) * kRP, nog * kRP, nig * kRP, 16 * h, 16 * h + 16, ig == 0);
      cur ^= 1;
    }
    // this lane's sums -> the warp's fp64 block (one owner per address: deterministic).  Rows / columns past the layer's
    // extent were computed on whatever rows follow in the buffer and are dropped here.
    if (active) {
      const int n_out = P.dims[l], n_in = P.dims[l - 1];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int o = og + nog * j;
        if (o >= n_out) continue;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const int c = ig + nig * i;
          if (c < n_in) atomicAdd(part + P.gw_off[l - 1] + o * n_in + c, (double)(acc[j][i].x + acc[j][i].y));
        }
        if (ig == 0) atomicAdd(part + P.gb_off[l - 1] + o, (double)bsum[j]);
      }
    }
  }
  // fold the CTA's warps into the first warp's block (fixed order)
  __syncthreads();
  {
    double* cta = P.part_dw + (size_t)blockIdx.x * nw * (size_t)P.n_params;
    for (int e = tid; e < P.n_params; e += nt) {
      double sacc = 0.0;
      for (int qq = 0; qq < nw; ++qq) sacc += __ldcg(cta + (size_t)qq * P.n_params + e);
      cta[e] = sacc;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

static bool shape_ok(const NetPlan& np, int* E, int* EC, int* G) {
  if (np.L != 7) return false;
  const int* dm = np.dims;
  static const int want_act[7] = {1, 1, 1, 0, 1, 1, 0};
  for (int l = 0; l < 7; ++l)
    if (np.act[l] != want_act[l]) return false;
  if (dm[0] != dm[7] || dm[1] != dm[2] || dm[2] != dm[3] || dm[5] != dm[6]) return false;
  *E = dm[1], *EC = dm[4], *G = dm[5];
  return *E == 20 && *EC >= 1 && *EC <= 3 && *G == 10 && dm[0] >= 1 && dm[0] <= 72;   // the instantiated shapes
}

static size_t main_smem(int img_floats, int d) { return ((size_t)img_floats + (size_t)d * kTile) * sizeof(float); }
static size_t dw_smem(int pairs, int rows_buf) { return (size_t)pairs * 2 * rows_buf * kRP * sizeof(float); }

// fills the plan; returns the workspace bytes it needs (carves `workspace` when it is not null)
static size_t make_plan(Plan* P, const NetPlan& np, long long B, void* workspace) {
  memset(P, 0, sizeof(*P));
  P->L = np.L, P->d = np.dims[0], P->drp = (int)round_up(np.dims[0], 12);
  int row = 0;
  for (int l = 0; l <= np.L; ++l) P->dims[l] = np.dims[l];
  for (int l = 0; l < np.L; ++l) P->aoff[l] = row, row += np.dims[l];
  for (int l = 1; l <= np.L; ++l) P->doff[l] = row, row += np.dims[l];
  P->n_rows = row;
  for (int l = 0; l < np.L; ++l) P->gw_off[l] = np.gw_off[l], P->gb_off[l] = np.gb_off[l];
  P->n_params = np.n_params;
  {
    const int ec = np.dims[4];
    P->img_floats = ec == 1 ? Img<20, 1, 10>::floats(P->d, P->drp) : ec == 2 ? Img<20, 2, 10>::floats(P->d, P->drp)
                                                                           : Img<20, 3, 10>::floats(P->d, P->drp);
  }
  // pack the layers into types: largest first, first type with room
  int order[kMaxLayers], lanes[kMaxLayers + 1];
  for (int l = 1; l <= np.L; ++l) lanes[l] = ((np.dims[l] + 3) / 4) * ((np.dims[l - 1] + 11) / 12), order[l - 1] = l;
  for (int i = 0; i < np.L; ++i)
    for (int j = i + 1; j < np.L; ++j)
      if (lanes[order[j]] > lanes[order[i]]) {
        const int t = order[i];
        order[i] = order[j], order[j] = t;
      }
  int used[kMaxTypes] = {0, 0, 0, 0}, rows[kMaxTypes] = {0, 0, 0, 0};
  P->n_types = 0;
  for (int i = 0; i < np.L; ++i) {
    const int l = order[i];
    int ty = 0;
    while (ty < P->n_types && used[ty] + lanes[l] > 32) ++ty;
    if (ty == P->n_types) {
      if (P->n_types == kMaxTypes || lanes[l] > 32) return 0;
      ++P->n_types;
    }
    const int s = P->type_n[ty]++;
    P->type_layer[ty][s] = l, P->type_lane0[ty][s] = used[ty];
    // four extra rows after each operand keep the out-of-extent rows a lane may touch inside the buffer
    P->type_xrow[ty][s] = rows[ty], rows[ty] += np.dims[l] + 4;
    P->type_zrow[ty][s] = rows[ty], rows[ty] += np.dims[l - 1] + 12;
    used[ty] += lanes[l];
  }
  P->rows_buf = 0;
  for (int ty = 0; ty < P->n_types; ++ty)
    if (rows[ty] > P->rows_buf) P->rows_buf = rows[ty];
  P->B = B, P->Bp = round_up(B, kTile);
  char* base = (char*)workspace;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  P->part_loss = (double*)take((size_t)sm_count() * 2 * 2 * sizeof(double));
  P->part_dw = (double*)take((size_t)sm_count() * 8 * P->n_params * sizeof(double));
  P->img = (float*)take((size_t)P->img_floats * sizeof(float));
  P->R = (float*)take((size_t)P->n_rows * P->Bp * sizeof(float));
  return off;
}

static int dw_pairs(int rows_buf) {
  for (int p = 4; p >= 1; --p)
    if (dw_smem(p, rows_buf) <= (size_t)max_smem_optin()) return p;
  return 0;
}

}  // namespace aefast

static std::atomic<int> g_ae_fast_mode{0};   // 0: use the fast kernels when the shape allows, 1: never

int fast_ae_set_mode(int mode) {
  if (mode != 0 && mode != 1) return CVF_E_ARG;
  g_ae_fast_mode = mode;
  return 0;
}

bool fast_ae_supported(const NetPlan& np) {
  int E, EC, G;
  if (g_ae_fast_mode != 0 || !aefast::shape_ok(np, &E, &EC, &G)) return false;
  aefast::Plan P;
  if (aefast::make_plan(&P, np, 512, nullptr) == 0) return false;
  return aefast::main_smem(P.img_floats, P.d) <= (size_t)max_smem_optin() && aefast::dw_pairs(P.rows_buf) > 0;
}

size_t fast_ae_workspace_bytes(const NetPlan& np, long long B) {
  aefast::Plan P;
  return aefast::make_plan(&P, np, B, nullptr);
}

int fast_ae_step(const NetPlan& np, const float* feat, const float* w, long long B, const float* params, double* sums_out,
                 double* grad_out, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  using namespace aefast;
  Plan P;
  const size_t need = make_plan(&P, np, B, workspace);
  if (need == 0 || need > ws_bytes) {
    set_error("workspace too small for the autoencoder fast path: %zu < %zu", ws_bytes, need);
    return CVF_E_WORKSPACE;
  }
  const int ec = np.dims[4];
#define CVF_AE_FAST_SHAPE(CALL) \
  do {                          \
    if (ec == 1) { CALL(20, 1, 10); } else if (ec == 2) { CALL(20, 2, 10); } else { CALL(20, 3, 10); } \
  } while (0)
#define CVF_AE_PACK(E_, EC_, G_) CVF_LAUNCH(K_FAST_PACK, stream, (pack_kernel<E_, EC_, G_><<<7, 256, 0, stream>>>(P, params)))
  CVF_AE_FAST_SHAPE(CVF_AE_PACK);
#undef CVF_AE_PACK
  CVF_CUDA(cudaGetLastError());
  {
    const size_t smem = (size_t)128 * (P.d | 1) * sizeof(float);
    CVF_CUDA(cudaFuncSetAttribute(prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long per_sm = (long long)(228 * 1024) / (long long)(smem + 1024);
    per_sm = per_sm < 1 ? 1 : per_sm > 8 ? 8 : per_sm;
    long long grid = (long long)sm_count() * per_sm;
    if (P.Bp / 128 < grid) grid = P.Bp / 128;
    CVF_LAUNCH(K_AE_FAST_PREP, stream, prep_kernel<<<(int)grid, 128, smem, stream>>>(P, feat));
    CVF_CUDA(cudaGetLastError());
  }
  int grid_main = 2 * sm_count();
  {
    const size_t smem = main_smem(P.img_floats, P.d);
    if (P.Bp / kTile < grid_main) grid_main = (int)(P.Bp / kTile);
#define CVF_AE_MAIN(E_, EC_, G_)                                                                                                  \
  do {                                                                                                                            \
    CVF_CUDA(cudaFuncSetAttribute(main_kernel<E_, EC_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    CVF_LAUNCH(K_AE_FAST_MAIN, stream, (main_kernel<E_, EC_, G_><<<grid_main, kThreads, smem, stream>>>(P, w, grad_out ? 1 : 0))); \
  } while (0)
    CVF_AE_FAST_SHAPE(CVF_AE_MAIN);
#undef CVF_AE_MAIN
#undef CVF_AE_FAST_SHAPE
    CVF_CUDA(cudaGetLastError());
  }
  CVF_LAUNCH(K_REDUCE, stream, reduce_partials_kernel<<<1, 32, 0, stream>>>(P.part_loss, grid_main, 2, 0, 2, sums_out));
  CVF_CUDA(cudaGetLastError());
  if (!grad_out) return 0;
  const int pairs = dw_pairs(P.rows_buf);
  const size_t smem = dw_smem(pairs, P.rows_buf);
  CVF_CUDA(cudaFuncSetAttribute(dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_items = (P.Bp / 32 + kRun - 1) / kRun * P.n_types;
  long long grid = sm_count();
  if ((n_items + pairs - 1) / pairs < grid) grid = (n_items + pairs - 1) / pairs;
  CVF_LAUNCH(K_AE_FAST_DW, stream, dw_kernel<<<(int)grid, 64 * pairs, smem, stream>>>(P));
  CVF_CUDA(cudaGetLastError());
  CVF_LAUNCH(K_REDUCE, stream,
             reduce_partials_kernel<<<(P.n_params + 127) / 128, 128, 0, stream>>>(P.part_dw, (int)grid, 2 * pairs * P.n_params, 0,
                                                                                 P.n_params, grad_out));
  CVF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace cvf
