// cvf_weights.cu -- device side of WeightedTrajectory's weight handling (reference utils.py:140-169): normalise the weights to
// mean 1, keep the states whose normalised weight lies in (min_w, max_w), renormalise the kept weights to mean 1, and return
// the (ascending) indices of the kept states so that the trajectory can be gathered on the device too.
//
// Five small launches on the caller's stream; every sum is fp64 with one owner per partial (deterministic):
//   block sums of w -> mean -> per-block (count, sum) of the kept states -> exclusive scan of the counts, kept mean ->
//   order-preserving compaction (block-local scan) writing idx_out and w_out.
#include "cvf_common.cuh"

namespace cvf {
namespace wsel {

constexpr int kThreads = 256, kPerThread = 8, kChunk = kThreads * kPerThread;

__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int q = 0; q < kThreads / 32; ++q) t += red[q];
  return t;
}

__global__ void __launch_bounds__(kThreads) sum_kernel(const double* __restrict__ w, long long n, double* __restrict__ part) {
  __shared__ double red[kThreads / 32];
  const long long base = (long long)blockIdx.x * kChunk;
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    const long long i = base + j * kThreads + threadIdx.x;
    if (i < n) acc += w[i];
  }
  const double t = block_sum(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

// scal[0] = mean of all weights
__global__ void __launch_bounds__(1024) mean_kernel(const double* __restrict__ part, int n_blocks, long long n, double* __restrict__ scal) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int b = threadIdx.x; b < n_blocks; b += 1024) acc += part[b];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < 32; ++q) t += red[q];
    scal[0] = t / (double)n;
  }
}

__device__ __forceinline__ bool kept(double w, double mean, double lo, double hi, double& wn) {
  wn = w / mean;
  return wn > lo && wn < hi;
}

__global__ void __launch_bounds__(kThreads) count_kernel(const double* __restrict__ w, long long n, const double* __restrict__ scal, double lo,
                                                          double hi, long long* __restrict__ cnt, double* __restrict__ ksum) {
  __shared__ double red[kThreads / 32];
  const long long base = (long long)blockIdx.x * kChunk;
  const double mean = scal[0];
  double acc = 0.0, c = 0.0;
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    const long long i = base + j * kThreads + threadIdx.x;
    double wn;
    if (i < n && kept(w[i], mean, lo, hi, wn)) acc += wn, c += 1.0;
  }
  const double ts = block_sum(acc, red);
  const double tc = block_sum(c, red);
  if (threadIdx.x == 0) ksum[blockIdx.x] = ts, cnt[blockIdx.x] = (long long)tc;
}

// offs[b] = number of kept states before block b; scal[1] = mean of the kept normalised weights; n_keep_out = total
__global__ void scan_kernel(const long long* __restrict__ cnt, const double* __restrict__ ksum, int n_blocks, long long* __restrict__ offs,
                            double* __restrict__ scal, long long* __restrict__ n_keep_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  long long run = 0;
  double s = 0.0;
  for (int b = 0; b < n_blocks; ++b) offs[b] = run, run += cnt[b], s += ksum[b];
  *n_keep_out = run;
  scal[1] = run > 0 ? s / (double)run : 1.0;
}

__global__ void __launch_bounds__(kThreads) compact_kernel(const double* __restrict__ w, long long n, const double* __restrict__ scal, double lo,
                                                            double hi, const long long* __restrict__ offs, long long* __restrict__ idx_out,
                                                            double* __restrict__ w_out) {
  __shared__ int wsum[kThreads / 32];
  const long long base = (long long)blockIdx.x * kChunk;
  const double mean = scal[0], kmean = scal[1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread t owns the kPerThread consecutive states base + t * kPerThread + j: positions stay in ascending order
  double wn[kPerThread];
  bool k[kPerThread];
  int mine = 0;
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    const long long i = base + (long long)threadIdx.x * kPerThread + j;
    k[j] = i < n && kept(w[i], mean, lo, hi, wn[j]);
    mine += k[j] ? 1 : 0;
  }
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int before = incl - mine;
  for (int q = 0; q < warp; ++q) before += wsum[q];
  long long pos = offs[blockIdx.x] + before;
#pragma unroll
  for (int j = 0; j < kPerThread; ++j)
    if (k[j]) {
      idx_out[pos] = base + (long long)threadIdx.x * kPerThread + j;
      w_out[pos] = wn[j] / kmean;
      ++pos;
    }
}

}  // namespace wsel
}  // namespace cvf

using namespace cvf;

extern "C" size_t cvf_weights_filter_workspace_bytes(int64_t n) {
  if (n < 1) return 0;
  const long long blocks = (n + wsel::kChunk - 1) / wsel::kChunk;
  return (size_t)blocks * (2 * sizeof(double) + 2 * sizeof(long long)) + 64;
}

extern "C" int cvf_weights_filter(const double* w, int64_t n, double min_w, double max_w, int64_t* idx_out, double* w_out,
                                  int64_t* n_keep_out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!w || !idx_out || !w_out || !n_keep_out || !workspace || n < 1) {
    set_error("cvf_weights_filter: null pointer or empty input");
    return CVF_E_ARG;
  }
  if (n > (1LL << 40)) {
    set_error("cvf_weights_filter: %lld weights", (long long)n);
    return CVF_E_UNSUPPORTED;
  }
  if (workspace_bytes < cvf_weights_filter_workspace_bytes(n)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, cvf_weights_filter_workspace_bytes(n));
    return CVF_E_WORKSPACE;
  }
  const int blocks = (int)((n + wsel::kChunk - 1) / wsel::kChunk);
  double* scal = (double*)workspace;                 // mean, kept mean (+ padding to 64 bytes)
  double* part = scal + 8;                           // [blocks] block sums, then the kept sums
  double* ksum = part + blocks;
  long long* cnt = (long long*)(ksum + blocks);
  long long* offs = cnt + blocks;
  CVF_LAUNCH(K_WEIGHTS, stream, wsel::sum_kernel<<<blocks, wsel::kThreads, 0, stream>>>(w, n, part));
  CVF_LAUNCH(K_WEIGHTS, stream, wsel::mean_kernel<<<1, 1024, 0, stream>>>(part, blocks, n, scal));
  CVF_LAUNCH(K_WEIGHTS, stream, wsel::count_kernel<<<blocks, wsel::kThreads, 0, stream>>>(w, n, scal, min_w, max_w, cnt, ksum));
  CVF_LAUNCH(K_WEIGHTS, stream, wsel::scan_kernel<<<1, 32, 0, stream>>>(cnt, ksum, blocks, offs, scal, (long long*)n_keep_out));
  CVF_LAUNCH(K_WEIGHTS, stream,
             wsel::compact_kernel<<<blocks, wsel::kThreads, 0, stream>>>(w, n, scal, min_w, max_w, offs, (long long*)idx_out, w_out));
  CVF_CUDA(cudaGetLastError());
  return 0;
}
