// mv_probe.cu -- micro-benchmarks behind the step-kernel design ("inner loops"): how many fp32 FMAs per
// clock per SM the candidate register-tile shapes of the small dense layers sustain on a B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mv_probe mv_probe.cu && ./mv_probe
#include <cuda_runtime.h>
#include <stdio.h>

typedef unsigned long long u64;
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  u64 rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)), "l"(*reinterpret_cast<u64*>(&c)));
  return *reinterpret_cast<float2*>(&rd);
}

// 0: register-only scalar FFMA chains      1: register-only FFMA2 chains
template <int MODE>
__global__ void __launch_bounds__(256) reg_probe(float* sink, int iters) {
  float2 a[16];
  const float2 x = make_float2(1.0f + 1e-7f * threadIdx.x, 1.0f - 1e-7f * threadIdx.x), y = make_float2(1e-9f * blockIdx.x, 1e-9f);
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = make_float2((float)i, (float)-i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, x.x, y.x); a[i].y = fmaf(a[i].y, x.y, y.y); }
      else a[i] = ffma2(a[i], x, y);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
  if (s == 12345.678f) sink[0] = s;
}

// Thread-private dense layers: H x H weights in shared memory, activations in registers, FPT frames per thread.
//  MODE 0: scalar FFMA, weights [o][k] read as LDS.128 along k (broadcast)
//  MODE 2: scalar FFMA, weights [k][o] read as LDS.128 along o (4 independent accumulators per load)
//  MODE 1: FFMA2 packed over output pairs, weights [k][o] read as LDS.128 along o (broadcast), activation duplicated
template <int MODE, int FPT, int H, int T>
__global__ void __launch_bounds__(T, 1) mv_probe(const float* __restrict__ wg, float* sink, int iters) {
  __shared__ __align__(16) float Wall[3 * H * H];
  for (int i = threadIdx.x; i < 3 * H * H; i += blockDim.x) Wall[i] = wg[i];
  __syncthreads();
  float x[FPT][H];
#pragma unroll
  for (int f = 0; f < FPT; ++f)
#pragma unroll
    for (int i = 0; i < H; ++i) x[f][i] = 1e-3f * (threadIdx.x + i + f);
  for (int it = 0; it < iters; ++it) {
    const float* W = Wall + (it % 3) * H * H;   // a different layer every iteration: the loads cannot be hoisted
    if (MODE == 0) {
      float y[FPT][H];
#pragma unroll
      for (int o = 0; o < H; ++o) {
#pragma unroll
        for (int f = 0; f < FPT; ++f) y[f][o] = 0.f;
#pragma unroll
        for (int k = 0; k < H; k += 4) {
          const float4 w = *reinterpret_cast<const float4*>(&W[o * H + k]);
#pragma unroll
          for (int f = 0; f < FPT; ++f) {
            y[f][o] = fmaf(w.x, x[f][k], y[f][o]);
            y[f][o] = fmaf(w.y, x[f][k + 1], y[f][o]);
            y[f][o] = fmaf(w.z, x[f][k + 2], y[f][o]);
            y[f][o] = fmaf(w.w, x[f][k + 3], y[f][o]);
          }
        }
      }
#pragma unroll
      for (int f = 0; f < FPT; ++f)
#pragma unroll
        for (int i = 0; i < H; ++i) x[f][i] = y[f][i] * 0.05f;
    } else if (MODE == 2) {
      float y[FPT][H];
#pragma unroll
      for (int f = 0; f < FPT; ++f)
#pragma unroll
        for (int o = 0; o < H; ++o) y[f][o] = 0.f;
#pragma unroll
      for (int k = 0; k < H; ++k) {
#pragma unroll
        for (int o = 0; o < H; o += 4) {
          const float4 w = *reinterpret_cast<const float4*>(&W[k * H + o]);
#pragma unroll
          for (int f = 0; f < FPT; ++f) {
            y[f][o] = fmaf(w.x, x[f][k], y[f][o]);
            y[f][o + 1] = fmaf(w.y, x[f][k], y[f][o + 1]);
            y[f][o + 2] = fmaf(w.z, x[f][k], y[f][o + 2]);
            y[f][o + 3] = fmaf(w.w, x[f][k], y[f][o + 3]);
          }
        }
      }
#pragma unroll
      for (int f = 0; f < FPT; ++f)
#pragma unroll
        for (int i = 0; i < H; ++i) x[f][i] = y[f][i] * 0.05f;
    } else {
      float2 y[FPT][H / 2];
#pragma unroll
      for (int f = 0; f < FPT; ++f)
#pragma unroll
        for (int o = 0; o < H / 2; ++o) y[f][o] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < H; ++k) {
#pragma unroll
        for (int o = 0; o < H; o += 4) {
          const float4 w = *reinterpret_cast<const float4*>(&W[k * H + o]);
#pragma unroll
          for (int f = 0; f < FPT; ++f) {
            const float2 xx = make_float2(x[f][k], x[f][k]);
            y[f][o / 2] = ffma2(make_float2(w.x, w.y), xx, y[f][o / 2]);
            y[f][o / 2 + 1] = ffma2(make_float2(w.z, w.w), xx, y[f][o / 2 + 1]);
          }
        }
      }
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
  float *sink, *w;
  cudaMalloc(&sink, 64);
  cudaMalloc(&w, 4096 * 4);
  cudaMemset(w, 0, 4096 * 4);
  printf("%s, %d SMs, max clock %.0f MHz (FMA/clk/SM figures assume the max clock)\n", p.name, sms, khz / 1e3);
  const double clk = khz * 1e3;
  {
    const int it = 20000;
    float ms = time_ms([&] { reg_probe<0><<<sms * 8, 256>>>(sink, it); });
    printf("reg scalar FFMA : %7.1f FMA/clk/SM  %.1f TFLOP/s\n", 32.0 * it * 256 * 8 / (ms * 1e-3 * clk), 2 * 32.0 * it * 256 * 8 * sms / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { reg_probe<1><<<sms * 8, 256>>>(sink, it); });
    printf("reg FFMA2       : %7.1f FMA/clk/SM  %.1f TFLOP/s\n", 32.0 * it * 256 * 8 / (ms * 1e-3 * clk), 2 * 32.0 * it * 256 * 8 * sms / (ms * 1e-3) / 1e12);
  }
#define RUN(MODE, FPT, T)                                                                                         \
  {                                                                                                               \
    const int it = 4000;                                                                                          \
    float ms = time_ms([&] { mv_probe<MODE, FPT, 20, T><<<sms, T>>>(w, sink, it); });                                \
    printf("mv mode %d fpt %d threads %3d: %7.1f FMA/clk/SM\n", MODE, FPT, T, 400.0 * FPT * it * T / (ms * 1e-3 * clk)); \
  }
  RUN(0, 1, 128) RUN(0, 1, 256) RUN(0, 1, 512) RUN(0, 2, 128) RUN(0, 2, 256) RUN(0, 2, 512) RUN(0, 4, 128) RUN(0, 4, 256)
  RUN(2, 1, 128) RUN(2, 1, 256) RUN(2, 1, 512) RUN(2, 2, 128) RUN(2, 2, 256) RUN(2, 2, 512) RUN(2, 4, 128) RUN(2, 4, 256)
  RUN(1, 1, 128) RUN(1, 1, 256) RUN(1, 1, 512) RUN(1, 2, 128) RUN(1, 2, 256) RUN(1, 2, 512) RUN(1, 4, 128) RUN(1, 4, 256)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
