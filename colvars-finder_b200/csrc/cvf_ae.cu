// cvf_ae.cu -- AutoEncoderTask.weighted_MSE_loss + backward on sm_100a.
//
// Replaces, per mini-batch (reference file:line under /root/reference/colvarsfinder):
//   core.py:664      out = model(X) = decoder(encoder(X))  (nn.py:114)  -> forward phases over the enc+dec chain
//   core.py:666      (weight * ((out-X)**2).sum(1)).sum() / weight.sum() -> fp64 sums {sum w|e|^2, sum w}
//   core.py:708      loss.backward()                                     -> reverse sweep + outer-product phases
// The division by sum(w) happens after the cross-GPU sum, on the caller's side.
//
// Same row engine as the eigenfunction kernel (cvf_common.cuh): activations of every layer stay in
// shared memory as [unit][frame] rows, weights stay resident, one persistent CTA per SM.
#include <string.h>

#include "cvf_common.cuh"

namespace cvf {

// cvf_ae_fast.cu: thread-private kernels for the notebook-sized chain  [d,20,20,20,2] + [2,10,10,d]
bool fast_ae_supported(const NetPlan& np);
size_t fast_ae_workspace_bytes(const NetPlan& np, long long B);
int fast_ae_step(const NetPlan& np, const float* feat, const float* w, long long B, const float* params, double* sums_out,
                 double* grad_out, void* workspace, size_t ws_bytes, cudaStream_t stream);
int fast_ae_set_mode(int mode);

struct AePlan {
  NetPlan net;
  int F, FS, FB, nthreads, ctas_per_sm;
  int a_row[kMaxLayers + 1];   // activations A_0 (input) .. A_L (output)
  int s_row[kMaxLayers + 1];   // adjoints of z_l, l = 1..L
  int row_w, row_err;
  int n_rows;
  int off_params, off_rows, off_red;
  size_t smem_bytes;
};

static size_t ae_layout(AePlan* P, int F) {
  const NetPlan& np = P->net;
  P->F = F, P->FS = F + 4, P->FB = F / 4;
  int r = 0;
  P->row_w = r++;
  P->row_err = r++;
  for (int l = 0; l <= np.L; ++l) P->a_row[l] = r, r += np.dims[l];
  for (int l = 1; l <= np.L; ++l) P->s_row[l] = r, r += np.dims[l];
  P->n_rows = r;
  int off = 0;
  P->off_params = off, off += np.smem_floats;
  P->off_red = off, off += 2 * 2 * 4;
  off = (off + 3) & ~3;
  P->off_rows = off, off += r * P->FS;
  P->smem_bytes = (size_t)off * sizeof(float);
  return P->smem_bytes;
}

template <bool GRAD, int FPL>
__global__ void __launch_bounds__(384, 1)
ae_kernel(const AePlan P, const float* __restrict__ feat, const float* __restrict__ target, const float* __restrict__ w, long long B,
          const float* __restrict__ params, double* __restrict__ partial) {
  extern __shared__ __align__(16) float smem[];
  const NetPlan& np = P.net;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int FS = P.FS, F = P.F, L = np.L;
  const int FB = F / FPL;
  typedef FVec<FPL> V;
  float* Wsm = smem + P.off_params;
  float* rows = smem + P.off_rows;
  double* red = reinterpret_cast<double*>(smem + P.off_red);
  load_net_params(np, params, Wsm, tid, nt);
  const int n_part = 2 + (GRAD ? np.n_params : 0);
  double* part = partial + (size_t)blockIdx.x * n_part;
  for (int i = tid; i < n_part; i += nt) part[i] = 0.0;
  double stat_acc = 0.0;   // thread 0: sum w |e|^2, thread 1: sum w
  __syncthreads();
  const int d0 = np.dims[0], dL = np.dims[L];
  const long long n_tiles = (B + F - 1) / F;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long f_base = tile * F;
    {
      // all loads of a thread are issued before its first store (8 in flight); next tile prefetched into L2
      const int total = F * d0;
      const long long last = B * (long long)d0 - 1;
      for (int base = 0; base < total; base += 8 * nt) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = base + j * nt + tid;
          if (idx < total) v[j] = __ldg(feat + min(f_base * d0 + idx, last));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int idx = base + j * nt + tid;
          if (idx < total) {
            const int f = idx / d0, u = idx - f * d0;
            rows[(P.a_row[0] + u) * FS + f] = v[j];
          }
        }
      }
      if (tid == 0) {
        const long long nxt = tile + gridDim.x;
        if (nxt < n_tiles) {
          const long long nb = nxt * F;
          const long long nf = min((long long)F, B - nb);
          const float* pn = feat + (size_t)nb * d0;
          const unsigned bytes = (unsigned)((nf * d0 * 4) & ~15LL);
          if (((reinterpret_cast<uintptr_t>(pn) & 15) == 0) && bytes > 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pn), "r"(bytes) : "memory");
        }
      }
    }
    for (int f = tid; f < F; f += nt) rows[P.row_w * FS + f] = f_base + f < B ? w[f_base + f] : 0.0f;
    __syncthreads();
    // ---- forward through encoder and decoder (nn.py:52-57,114)
    for (int l = 0; l < L; ++l) {
      const int nin = np.dims[l], nout = np.dims[l + 1];
      const float* in = rows + P.a_row[l] * FS;
      float* out = rows + P.a_row[l + 1] * FS;
      const int act = np.act[l];
      const int items = ((nout + 3) >> 2) * FB;
      for (int it = tid; it < items; it += nt) {
        const int ob = it / FB, fb = it - ob * FB;
        float acc[4][FPL];
        tile_fwd<FPL>(acc, Wsm + np.w_off[l], np.ld[l], 4 * ob, nout, nin, in, FS, FPL * fb);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int o = 4 * ob + j;
          if (o < nout) {
            const float b = Wsm[np.b_off[l] + o];
            V v;
#pragma unroll
            for (int f = 0; f < FPL; ++f) v.v[f] = cvf_act(act, acc[j][f] + b);
            v.st(out + o * FS + FPL * fb);
          }
        }
      }
      __syncthreads();
    }
    // ---- per-frame squared error, batch sums, and the seed s_L = 2 w (out - X)   (core.py:666)
    if (tid < F) {
      const int f = tid;
      const float wf = rows[P.row_w * FS + f];
      const float* X = rows + P.a_row[0] * FS;
      const float* O = rows + P.a_row[L] * FS;
      float* S = rows + P.s_row[L] * FS;
      double e = 0.0;
      // reconstruction target: the input itself, or (time-lagged autoencoder, RegAutoEncoderTask.weighted_MSE_loss,
      // core.py:876-887) the features of another frame, read row-wise from global memory
      const float* T = target != nullptr && f_base + f < B ? target + (size_t)(f_base + f) * dL : nullptr;
      for (int o = 0; o < dL; ++o) {
        const float d = O[o * FS + f] - (T ? __ldg(T + o) : X[o * FS + f]);
        e += (double)d * (double)d;
        if (GRAD) S[o * FS + f] = 2.0f * wf * d;
      }
      double v0 = (double)wf * e, v1 = (double)wf;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      }
      if ((tid & 31) == 0) red[(tid >> 5)] = v0, red[4 + (tid >> 5)] = v1;
    }
    __syncthreads();
    if (tid < 2) {
      double v = 0.0;
      for (int q = 0; q < F / 32; ++q) v += red[tid * 4 + q];
      stat_acc += v;
    }
    if (GRAD) {
      // ---- reverse sweep: dW_l += s_l (x) A_{l-1}, db_l += sum_f s_l, s_{l-1} = (W_l^T s_l) .* act'(A_{l-1})
      for (int l = L - 1; l >= 0; --l) {
        const int nin = np.dims[l], nout = np.dims[l + 1];
        const float* S = rows + P.s_row[l + 1] * FS;
        const float* Ain = rows + P.a_row[l] * FS;
        // both kinds of work only read s_l and A_{l-1}: outer products first, then the back-propagation items
        outer_layer(0, tid, nt, nout, nin, S, Ain, nullptr, nullptr, FS, F, part + 2 + np.gw_off[l], part + 2 + np.gb_off[l]);
        const int n_back = l > 0 ? ((nin + 3) >> 2) * FB : 0;
        for (int it2 = tid; it2 < n_back; it2 += nt) {
          const int ib = it2 / FB, fb = it2 - ib * FB;
          float acc[4][FPL];
          tile_tr<FPL>(acc, Wsm + np.w_off[l], np.ld[l], 4 * ib, nout, S, FS, FPL * fb);
          float* Sout = rows + P.s_row[l] * FS;
          const int act = np.act[l - 1];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = 4 * ib + j;
            if (i < nin) {
              V v;
              if (act) {
                const V a = V::ld(Ain + i * FS + FPL * fb);
#pragma unroll
                for (int f = 0; f < FPL; ++f) v.v[f] = acc[j][f] * cvf_act_d1(act, a.v[f]);
              } else {
#pragma unroll
                for (int f = 0; f < FPL; ++f) v.v[f] = acc[j][f];
              }
              v.st(Sout + i * FS + FPL * fb);
            }
          }
        }
        __syncthreads();
      }
    } else {
      __syncthreads();
    }
  }
  if (tid < 2) part[tid] = stat_acc;
}

// cvf_ae_wide.cu: layer-by-layer dense products for networks whose weights do not fit shared memory
int wide_ae_set_mode(int mode);
size_t wide_ae_workspace_bytes(const NetPlan& np, long long B);
int wide_ae_step(const NetPlan& np, const float* feat, const float* w, long long B, const float* params, double* sums_out,
                 double* grad_out, void* workspace, size_t ws_bytes, cudaStream_t stream);

}  // namespace cvf

using namespace cvf;

static int ae_plan(const cvf_mlp* net, AePlan* P) {
  memset(P, 0, sizeof(*P));
  int e = make_net_plan(net, &P->net);
  if (e) return e;
  if (P->net.dims[0] != P->net.dims[P->net.L]) {
    set_error("autoencoder chain must map R^d to R^d (got %d -> %d)", P->net.dims[0], P->net.dims[P->net.L]);
    return CVF_E_ARG;
  }
  const size_t cap = (size_t)max_smem_optin();
  static const int Fs[3] = {128, 64, 32};
  int c = 0;
  for (; c < 3; ++c) {
    if (ae_layout(P, Fs[c]) <= cap) break;
    if (c == 2) {
      set_error("autoencoder state does not fit shared memory (%zu B at 32 frames, %zu available)", P->smem_bytes, cap);
      return CVF_E_UNSUPPORTED;
    }
  }
  P->nthreads = 384;
  P->ctas_per_sm = 1;
  // Two CTAs of half the frames and half the threads per SM instead of one: the phases of the row engine end in CTA-wide
  // barriers, and a second, independent CTA fills the issue slots the first one leaves while it waits at them.
  if (c < 2 && 2 * (ae_layout(P, Fs[c + 1]) + 1024) <= (size_t)228 * 1024) {
    P->nthreads = 192;
    P->ctas_per_sm = 2;
  } else {
    ae_layout(P, Fs[c]);
  }
  return 0;
}

// 1 if the fused shared-memory kernel takes this chain, 0 if the layer-wise products do, < 0 if neither can
static int ae_path(const cvf_mlp* net, AePlan* P) {
  const int e = ae_plan(net, P);
  if (e == 0) return 1;
  if (e != CVF_E_UNSUPPORTED) return e;
  NetPlan np;
  if (make_net_plan(net, &np)) return CVF_E_ARG;
  if (np.dims[0] != np.dims[np.L] || np.act[np.L - 1] != 0) {
    set_error("layer-wise autoencoder path: the chain must map R^d to R^d and end in a linear layer");
    return CVF_E_UNSUPPORTED;
  }
  for (int l = 0; l < np.L; ++l)
    if (np.act[l] != CVF_ACT_NONE && np.act[l] != CVF_ACT_TANH) {
      set_error("layer-wise autoencoder path (chains too wide for shared memory): tanh only");
      return CVF_E_UNSUPPORTED;
    }
  return 0;
}

extern "C" size_t cvf_ae_workspace_bytes(const cvf_mlp* net, int64_t B) {
  AePlan P;
  const int path = ae_path(net, &P);
  if (path < 0 || B < 1) return 0;
  if (path == 0) return wide_ae_workspace_bytes(P.net, B);
  const size_t general = (size_t)(2 + P.net.n_params) * sizeof(double) * (size_t)sm_count() * 2;
  if (fast_ae_supported(P.net)) {
    const size_t fast = fast_ae_workspace_bytes(P.net, B);
    return fast > general ? fast : general;
  }
  return general;
}

static int ae_step_impl(const float* feat, const float* target, const float* w, int64_t B, const cvf_mlp* net, const float* params,
                        double* sums_out, double* grad_out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  AePlan P;
  const int path = ae_path(net, &P);
  if (path < 0) return path;
  if (!feat || !w || !params || !sums_out || !workspace || B < 1) {
    set_error("cvf_ae_step: null pointer or empty batch");
    return CVF_E_ARG;
  }
  if (target == feat) target = nullptr;
  if (target != nullptr && path == 0) {
    set_error("cvf_ae_step_target: a separate reconstruction target needs a chain that fits shared memory");
    return CVF_E_UNSUPPORTED;
  }
  if (target != nullptr && P.net.dims[P.net.L] != P.net.dims[0]) {
    set_error("cvf_ae_step_target: output width %d differs from the feature width %d", P.net.dims[P.net.L], P.net.dims[0]);
    return CVF_E_ARG;
  }
  if (path == 0) return wide_ae_step(P.net, feat, w, B, params, sums_out, grad_out, workspace, workspace_bytes, stream);
  if (target == nullptr && fast_ae_supported(P.net))
    return fast_ae_step(P.net, feat, w, B, params, sums_out, grad_out, workspace, workspace_bytes, stream);
  const bool grad = grad_out != nullptr;
  const int n_part = 2 + (grad ? P.net.n_params : 0);
  const long long n_tiles = (B + P.F - 1) / P.F;
  int grid = sm_count() * P.ctas_per_sm;
  if (n_tiles < grid) grid = (int)n_tiles;
  if ((size_t)grid * n_part * sizeof(double) > workspace_bytes) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, (size_t)grid * n_part * sizeof(double));
    return CVF_E_WORKSPACE;
  }
  if (grad) {
    CVF_CUDA(cudaFuncSetAttribute(ae_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem_bytes));
    CVF_LAUNCH(K_AE_STEP, stream, ae_kernel<true, 2><<<grid, P.nthreads, P.smem_bytes, stream>>>(P, feat, target, w, B, params, (double*)workspace));
  } else {
    CVF_CUDA(cudaFuncSetAttribute(ae_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem_bytes));
    CVF_LAUNCH(K_AE_STEP, stream, ae_kernel<false, 2><<<grid, P.nthreads, P.smem_bytes, stream>>>(P, feat, target, w, B, params, (double*)workspace));
  }
  CVF_CUDA(cudaGetLastError());
  // partial layout per CTA: [sum w|e|^2, sum w, grad...]; two reductions keep the output buffers separate
  CVF_LAUNCH(K_REDUCE, stream, reduce_partials_kernel<<<1, 32, 0, stream>>>((const double*)workspace, grid, n_part, 0, 2, sums_out));
  if (grad)
    CVF_LAUNCH(K_REDUCE, stream, reduce_partials_kernel<<<(P.net.n_params + 127) / 128, 128, 0, stream>>>((const double*)workspace, grid, n_part, 2,
                                                                             P.net.n_params, grad_out));
  CVF_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int cvf_ae_step(const float* feat, const float* w, int64_t B, const cvf_mlp* net, const float* params,
                           double* sums_out, double* grad_out, void* workspace, size_t workspace_bytes, void* stream) {
  return ae_step_impl(feat, nullptr, w, B, net, params, sums_out, grad_out, workspace, workspace_bytes, stream);
}

extern "C" int cvf_ae_step_target(const float* feat, const float* target, const float* w, int64_t B, const cvf_mlp* net,
                                  const float* params, double* sums_out, double* grad_out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  return ae_step_impl(feat, target, w, B, net, params, sums_out, grad_out, workspace, workspace_bytes, stream);
}

extern "C" int cvf_ae_set_wide_path(int32_t mode) {
  if (wide_ae_set_mode(mode)) {
    set_error("cvf_ae_set_wide_path: mode must be 0 (tensor cores) or 1 (fp32 SIMT)");
    return CVF_E_ARG;
  }
  return 0;
}

extern "C" int cvf_ae_set_fast_path(int32_t mode) {
  if (cvf::fast_ae_set_mode(mode)) {
    cvf::set_error("cvf_ae_set_fast_path: mode must be 0 (auto) or 1 (general kernels)");
    return CVF_E_ARG;
  }
  return 0;
}
