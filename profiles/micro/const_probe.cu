// constant-bank operand probe: thread-private dense layers with weights read as FFMA constant operands
#include <cuda_runtime.h>
#include <stdio.h>
__constant__ float CW[16000];
// NL layers of HxH cycled; FPT frames per thread
template <int FPT, int H, int T, int NL>
__global__ void __launch_bounds__(T, 1) cprobe(float* sink, int iters) {
  float x[FPT][H];
#pragma unroll
  for (int f = 0; f < FPT; ++f)
#pragma unroll
    for (int i = 0; i < H; ++i) x[f][i] = 1e-3f * (threadIdx.x + i + f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      float y[FPT][H];
#pragma unroll
      for (int f = 0; f < FPT; ++f)
#pragma unroll
        for (int o = 0; o < H; ++o) y[f][o] = 0.f;
#pragma unroll
      for (int k = 0; k < H; ++k)
#pragma unroll
        for (int o = 0; o < H; ++o)
#pragma unroll
          for (int f = 0; f < FPT; ++f) y[f][o] = fmaf(CW[l * H * H + k * H + o], x[f][k], y[f][o]);
#pragma unroll
      for (int f = 0; f < FPT; ++f)
#pragma unroll
        for (int i = 0; i < H; ++i) x[f][i] = y[f][i] * 0.05f;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int f = 0; f < FPT; ++f)
#pragma unroll
    for (int i = 0; i < H; ++i) s += x[f][i];
  if (s == 12345.678f) sink[0] = s;
}
typedef unsigned long long u64;
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  u64 rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)), "l"(*reinterpret_cast<u64*>(&c)));
  return *reinterpret_cast<float2*>(&rd);
}
template <int FPT, int H, int T, int NL>
__global__ void __launch_bounds__(T, 1) cprobe2(float* sink, int iters) {
  float x[FPT][H];
#pragma unroll
  for (int f = 0; f < FPT; ++f)
#pragma unroll
    for (int i = 0; i < H; ++i) x[f][i] = 1e-3f * (threadIdx.x + i + f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      float2 y[FPT][H / 2];
#pragma unroll
      for (int f = 0; f < FPT; ++f)
#pragma unroll
        for (int o = 0; o < H / 2; ++o) y[f][o] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < H; ++k)
#pragma unroll
        for (int o = 0; o < H; o += 2)
#pragma unroll
          for (int f = 0; f < FPT; ++f)
            y[f][o / 2] = ffma2(make_float2(CW[l * H * H + k * H + o], CW[l * H * H + k * H + o + 1]), make_float2(x[f][k], x[f][k]), y[f][o / 2]);
#pragma unroll
      for (int f = 0; f < FPT; ++f)
#pragma unroll
        for (int i = 0; i < H / 2; ++i) { x[f][2 * i] = y[f][i].x * 0.05f; x[f][2 * i + 1] = y[f][i].y * 0.05f; }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int f = 0; f < FPT; ++f)
#pragma unroll
    for (int i = 0; i < H; ++i) s += x[f][i];
  if (s == 12345.678f) sink[0] = s;
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
  float* sink;
  cudaMalloc(&sink, 64);
  const double clk = khz * 1e3;
  printf("%s %d SMs %.0f MHz\n", p.name, sms, khz / 1e3);
#define RUN(FPT, T, NL)                                                                                          \
  {                                                                                                              \
    const int it = 6000 / NL;                                                                                    \
    float ms = time_ms([&] { cprobe<FPT, 20, T, NL><<<sms, T>>>(sink, it); });                                   \
    printf("const fpt %d threads %4d layers %2d (%5d B): %7.1f FMA/clk/SM\n", FPT, T, NL, NL * 1600, 400.0 * NL * FPT * it * T / (ms * 1e-3 * clk)); \
  }
  RUN(1, 256, 1) RUN(1, 512, 1) RUN(1, 1024, 1)
  RUN(1, 256, 3) RUN(1, 512, 3) RUN(1, 1024, 3)
  RUN(1, 512, 6) RUN(1, 1024, 6)
  RUN(2, 256, 3) RUN(2, 512, 3)
  RUN(2, 256, 6) RUN(2, 512, 6)
#define RUN2(FPT, T, NL)                                                                                          \
  {                                                                                                              \
    const int it = 6000 / NL;                                                                                    \
    float ms = time_ms([&] { cprobe2<FPT, 20, T, NL><<<sms, T>>>(sink, it); });                                   \
    printf("const FFMA2 fpt %d threads %4d layers %2d (%5d B): %7.1f FMA/clk/SM\n", FPT, T, NL, NL * 1600, 400.0 * NL * FPT * it * T / (ms * 1e-3 * clk)); \
  }
  RUN2(1, 256, 3) RUN2(1, 512, 3) RUN2(1, 1024, 3) RUN2(2, 256, 3) RUN2(2, 512, 3) RUN2(1, 512, 6) RUN2(1, 1024, 6)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
