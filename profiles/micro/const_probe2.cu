// layer-1-like loop with runtime-indexed constant operands: footprint / unroll / thread-count sweep
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__constant__ float4 c_img[3968];
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  u64 rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)), "l"(*reinterpret_cast<u64*>(&c)));
  return *reinterpret_cast<float2*>(&rd);
}
template <int T, int UNR>
__global__ void __launch_bounds__(T, 1) kern(const float* __restrict__ Y, float* out, int d_r, long long Bp, int k, int img4, int iters) {
  const int tid = threadIdx.x;
  const float* Yf = Y + blockIdx.x * T + tid;
  float s = 0;
  for (int it = 0; it < iters; ++it)
  for (int n = 0; n < k; ++n) {
    const int wb = n * img4;
    float2 z[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) z[j] = make_float2(0.f, 0.f);
#pragma unroll UNR
    for (int kk = 0; kk < d_r; ++kk) {
      const float x = __ldg(Yf + (size_t)kk * Bp);
      const int wr = wb + kk * 5;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const float4 wv = c_img[wr + j];
        z[2 * j] = ffma2(make_float2(wv.x, wv.y), make_float2(x, x), z[2 * j]);
        z[2 * j + 1] = ffma2(make_float2(wv.z, wv.w), make_float2(x, x), z[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 10; ++j) s += z[j].x + z[j].y;
  }
  out[blockIdx.x * T + tid] = s;
}
// fully immediate version: 3 nets x 66 rows unrolled
template <int T>
__global__ void __launch_bounds__(T, 1) kern_imm(const float* __restrict__ Y, float* out, long long Bp, int iters) {
  const int tid = threadIdx.x;
  const float* Yf = Y + blockIdx.x * T + tid;
  float s = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    float2 z[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) z[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int kk = 0; kk < 66; ++kk) {
      const float x = __ldg(Yf + (size_t)kk * Bp);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const float4 wv = c_img[n * 1141 + kk * 5 + j];
        z[2 * j] = ffma2(make_float2(wv.x, wv.y), make_float2(x, x), z[2 * j]);
        z[2 * j + 1] = ffma2(make_float2(wv.z, wv.w), make_float2(x, x), z[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 10; ++j) s += z[j].x + z[j].y;
  }
  }
  out[blockIdx.x * T + tid] = s;
}
template <typename F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 3;
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;
  const long long Bp = (long long)sms * 1024;
  float *Y, *out;
  cudaMalloc(&Y, Bp * 72 * 4);
  cudaMemset(Y, 0, Bp * 72 * 4);
  cudaMalloc(&out, Bp * 4);
  const int iters = 40;
#define RUN(T, UNR, IMG4)                                                                                         \
  {                                                                                                               \
    float ms = time_ms([&] { kern<T, UNR><<<sms, T>>>(Y, out, 66, Bp, 3, IMG4, iters); });                          \
    printf("dyn  T %4d unroll %d img %5d B: %7.1f FMA/clk/SM\n", T, UNR, IMG4 * 16, 66.0 * 20 * 3 * iters * T / (ms * 1e-3 * clk)); \
  }
  RUN(768, 4, 1141) RUN(768, 2, 1141) RUN(768, 1, 1141) RUN(768, 2, 330) RUN(768, 2, 0) RUN(1024, 2, 1141) RUN(1024, 2, 0) RUN(512, 2, 1141) RUN(512, 2, 0)
  RUN(768, 8, 1141)
  {
    float ms = time_ms([&] { kern_imm<768><<<sms, 768>>>(Y, out, Bp, iters); });
    printf("imm  T  768: %7.1f FMA/clk/SM\n", 66.0 * 20 * 3 * iters * 768 / (ms * 1e-3 * clk));
    ms = time_ms([&] { kern_imm<1024><<<sms, 1024>>>(Y, out, Bp, iters); });
    printf("imm  T 1024: %7.1f FMA/clk/SM\n", 66.0 * 20 * 3 * iters * 1024 / (ms * 1e-3 * clk));
    ms = time_ms([&] { kern_imm<512><<<sms, 512>>>(Y, out, Bp, iters); });
    printf("imm  T  512: %7.1f FMA/clk/SM\n", 66.0 * 20 * 3 * iters * 512 / (ms * 1e-3 * clk));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
