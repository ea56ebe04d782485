// cvf_common.cuh -- error plumbing, launch helpers and the shared-memory "row engine" used by the
// eigenfunction and autoencoder step kernels.
//
// Data layout inside a CTA.  A tile is F frames (F = 32, 64 or 128).  Every per-frame quantity lives in
// shared memory as a ROW of F floats (row stride FS = F + 4 floats, so rows stay 16-byte aligned and rows
// r, r+1, .. r+7 start in different bank groups): row[unit][frame].  A lane owns 4 consecutive frames
// (one LDS.128 / STS.128 per row access).  All the small dense layers are evaluated as register-tiled
// products over these rows:
//     forward     out[o][f] = sum_i W[o][i] * in[i][f]      thread tile: 4 outputs x 4 frames
//     transposed  out[i][f] = sum_o W[o][i] * in[o][f]      thread tile: 4 inputs  x 4 frames
//     outer       dW[o][i] += sum_f X[o][f] * Z[i][f]       thread tile: 4 x 4 weights, loop over frames
// with the weights W resident in shared memory in torch's [out][in] layout, `in` padded to a multiple
// of 4 with zeros.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cvf.h"

namespace cvf {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int sm_count();
int max_smem_optin();

#define CVF_CUDA(call)                                   \
  do {                                                   \
    int _e = cvf::check_cuda((call), #call);             \
    if (_e) return _e;                                   \
  } while (0)

constexpr int kMaxLayers = CVF_MAX_LAYERS;
constexpr int kMaxK = CVF_MAX_K;

// Shape and shared-memory placement of one Linear+activation chain.
struct NetPlan {
  int L;
  int dims[kMaxLayers + 1];
  int act[kMaxLayers];
  int ld[kMaxLayers];      // padded input width of layer l (multiple of 4)
  int w_off[kMaxLayers];   // float offset of W_l inside the shared-memory parameter block of one net
  int b_off[kMaxLayers];
  int gw_off[kMaxLayers];  // float offset of W_l inside the caller's flat parameter vector (torch order)
  int gb_off[kMaxLayers];
  int n_params;            // per net, unpadded
  int smem_floats;         // per net, padded (multiple of 4)
};

inline int round4(int v) { return (v + 3) & ~3; }

inline int make_net_plan(const cvf_mlp* net, NetPlan* p) {
  if (!net || net->n_layers < 1 || net->n_layers > kMaxLayers) {
    set_error("cvf_mlp: n_layers must be in [1,%d]", kMaxLayers);
    return CVF_E_UNSUPPORTED;
  }
  p->L = net->n_layers;
  int so = 0, go = 0;
  for (int l = 0; l <= p->L; ++l) {
    if (net->dims[l] < 1) {
      set_error("cvf_mlp: dims[%d] = %d", l, net->dims[l]);
      return CVF_E_ARG;
    }
    p->dims[l] = net->dims[l];
  }
  for (int l = 0; l < p->L; ++l) {
    p->act[l] = net->act[l] ? 1 : 0;
    p->ld[l] = round4(p->dims[l]);
    p->w_off[l] = so;
    so += p->dims[l + 1] * p->ld[l];
    p->b_off[l] = so;
    so += round4(p->dims[l + 1]);
    p->gw_off[l] = go;
    go += p->dims[l + 1] * p->dims[l];
    p->gb_off[l] = go;
    go += p->dims[l + 1];
  }
  p->n_params = go;
  p->smem_floats = so;
  return 0;
}

#if defined(__CUDACC__)

// ---- device helpers --------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// copy one net's parameters from the caller's flat vector into the padded shared-memory block
__device__ inline void load_net_params(const NetPlan& np, const float* __restrict__ g, float* s, int tid, int nt) {
  for (int i = tid; i < np.smem_floats; i += nt) s[i] = 0.0f;
  __syncthreads();
  for (int l = 0; l < np.L; ++l) {
    const int nin = np.dims[l], nout = np.dims[l + 1];
    for (int i = tid; i < nin * nout; i += nt) s[np.w_off[l] + (i / nin) * np.ld[l] + (i % nin)] = g[np.gw_off[l] + i];
    for (int i = tid; i < nout; i += nt) s[np.b_off[l] + i] = g[np.gb_off[l] + i];
  }
}

// acc[j][f] = sum_i W[o0+j][i] * in[i][f0+f]      (rows of W clamped to n_out-1 for the tail block)
__device__ __forceinline__ void tile_fwd(float (&acc)[4][4], const float* __restrict__ W, int ld, int o0, int n_out,
                                         int n_in, const float* __restrict__ in, int FS, int f0) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int f = 0; f < 4; ++f) acc[j][f] = 0.0f;
  const float* w0 = W + min(o0 + 0, n_out - 1) * ld;
  const float* w1 = W + min(o0 + 1, n_out - 1) * ld;
  const float* w2 = W + min(o0 + 2, n_out - 1) * ld;
  const float* w3 = W + min(o0 + 3, n_out - 1) * ld;
  const float* a = in + f0;
  int i = 0;
#pragma unroll 2
  for (; i + 4 <= n_in; i += 4) {
    const float4 wv[4] = {ld4(w0 + i), ld4(w1 + i), ld4(w2 + i), ld4(w3 + i)};
    const float4 a0 = ld4(a + (i + 0) * FS), a1 = ld4(a + (i + 1) * FS), a2 = ld4(a + (i + 2) * FS), a3 = ld4(a + (i + 3) * FS);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j][0] = fmaf(wv[j].x, a0.x, acc[j][0]); acc[j][1] = fmaf(wv[j].x, a0.y, acc[j][1]);
      acc[j][2] = fmaf(wv[j].x, a0.z, acc[j][2]); acc[j][3] = fmaf(wv[j].x, a0.w, acc[j][3]);
      acc[j][0] = fmaf(wv[j].y, a1.x, acc[j][0]); acc[j][1] = fmaf(wv[j].y, a1.y, acc[j][1]);
      acc[j][2] = fmaf(wv[j].y, a1.z, acc[j][2]); acc[j][3] = fmaf(wv[j].y, a1.w, acc[j][3]);
      acc[j][0] = fmaf(wv[j].z, a2.x, acc[j][0]); acc[j][1] = fmaf(wv[j].z, a2.y, acc[j][1]);
      acc[j][2] = fmaf(wv[j].z, a2.z, acc[j][2]); acc[j][3] = fmaf(wv[j].z, a2.w, acc[j][3]);
      acc[j][0] = fmaf(wv[j].w, a3.x, acc[j][0]); acc[j][1] = fmaf(wv[j].w, a3.y, acc[j][1]);
      acc[j][2] = fmaf(wv[j].w, a3.z, acc[j][2]); acc[j][3] = fmaf(wv[j].w, a3.w, acc[j][3]);
    }
  }
  for (; i < n_in; ++i) {
    const float4 av = ld4(a + i * FS);
    const float wj[4] = {w0[i], w1[i], w2[i], w3[i]};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j][0] = fmaf(wj[j], av.x, acc[j][0]); acc[j][1] = fmaf(wj[j], av.y, acc[j][1]);
      acc[j][2] = fmaf(wj[j], av.z, acc[j][2]); acc[j][3] = fmaf(wj[j], av.w, acc[j][3]);
    }
  }
}

// acc[j][f] = sum_o W[o][i0+j] * in[o][f0+f]
__device__ __forceinline__ void tile_tr(float (&acc)[4][4], const float* __restrict__ W, int ld, int i0, int n_out,
                                        const float* __restrict__ in, int FS, int f0) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int f = 0; f < 4; ++f) acc[j][f] = 0.0f;
  const float* wp = W + i0;
  const float* a = in + f0;
#pragma unroll 4
  for (int o = 0; o < n_out; ++o) {
    const float4 wv = ld4(wp + o * ld);
    const float4 av = ld4(a + o * FS);
    acc[0][0] = fmaf(wv.x, av.x, acc[0][0]); acc[0][1] = fmaf(wv.x, av.y, acc[0][1]);
    acc[0][2] = fmaf(wv.x, av.z, acc[0][2]); acc[0][3] = fmaf(wv.x, av.w, acc[0][3]);
    acc[1][0] = fmaf(wv.y, av.x, acc[1][0]); acc[1][1] = fmaf(wv.y, av.y, acc[1][1]);
    acc[1][2] = fmaf(wv.y, av.z, acc[1][2]); acc[1][3] = fmaf(wv.y, av.w, acc[1][3]);
    acc[2][0] = fmaf(wv.z, av.x, acc[2][0]); acc[2][1] = fmaf(wv.z, av.y, acc[2][1]);
    acc[2][2] = fmaf(wv.z, av.z, acc[2][2]); acc[2][3] = fmaf(wv.z, av.w, acc[2][3]);
    acc[3][0] = fmaf(wv.w, av.x, acc[3][0]); acc[3][1] = fmaf(wv.w, av.y, acc[3][1]);
    acc[3][2] = fmaf(wv.w, av.z, acc[3][2]); acc[3][3] = fmaf(wv.w, av.w, acc[3][3]);
  }
}

__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

// Outer-product accumulation for one layer:  dW[o][i] += sum_f X1[o][f] Z1[i][f] (+ X2[o][f] Z2[i][f]),
// db[o] += sum_f X1[o][f].  Work item = 4 strided outputs x 4 strided inputs; results are added to the
// CTA's fp64 partial vector `part` (torch parameter order).  Items [item0, item0+n_items) of this layer
// are spread over the calling threads by the caller.
__device__ __forceinline__ void outer_item(int item, int n_out, int n_in, const float* __restrict__ X1,
                                           const float* __restrict__ Z1, const float* __restrict__ X2,
                                           const float* __restrict__ Z2, int FS, int F, double* __restrict__ part_w,
                                           double* __restrict__ part_b) {
  const int nob = (n_out + 3) >> 2, nib = (n_in + 3) >> 2;
  const int ob = item / nib, ib = item - ob * nib;
  int orow[4], irow[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    orow[j] = min(ob + j * nob, n_out - 1);
    irow[j] = min(ib + j * nib, n_in - 1);
  }
  float acc[4][4];
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
  for (int f = 0; f < F; f += 4) {
    float4 x[4], z[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x[j] = ld4(X1 + orow[j] * FS + f);
      z[j] = ld4(Z1 + irow[j] * FS + f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bacc[j] += (x[j].x + x[j].y) + (x[j].z + x[j].w);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] += dot4(x[j], z[i]);
    }
    if (X2 != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        x[j] = ld4(X2 + orow[j] * FS + f);
        z[j] = ld4(Z2 + irow[j] * FS + f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] += dot4(x[j], z[i]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int o = ob + j * nob;
    if (o < n_out) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ii = ib + i * nib;
        if (ii < n_in) part_w[o * n_in + ii] += (double)acc[j][i];
      }
      if (ib == 0 && part_b != nullptr) part_b[o] += (double)bacc[j];
    }
  }
}

// deterministic sum of per-CTA fp64 partial vectors: out[i] = sum_b part[b * stride + off + i], i < n
static __global__ void reduce_partials_kernel(const double* __restrict__ part, int n_blocks, int stride, int off, int n,
                                              double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int b = 0; b < n_blocks; ++b) s += part[(size_t)b * stride + off + i];
  out[i] = s;
}

#endif  // __CUDACC__

}  // namespace cvf
