// outer_probe.cu -- how fast does the warp-level outer product of pass 2b run when its operand rows are already in
// shared memory?  (profiles/micro: design measurements, not product code.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o outer_probe outer_probe.cu && ./outer_probe
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<u64*>(&a)), "l"(*reinterpret_cast<u64*>(&b)), "l"(*reinterpret_cast<u64*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
constexpr int RP = 36;

template <int TO, int TI, int UNROLL>
__device__ __forceinline__ void outer_tile(float2 (&acc)[TO][TI], const float* Xa, const float* Za, const float* Xb, const float* Zb,
                                           int strideX, int strideZ) {
#pragma unroll UNROLL
  for (int f = 0; f < 32; f += 4) {
    float4 x[TO], z[TI];
#pragma unroll
    for (int j = 0; j < TO; ++j) x[j] = ld4(Xa + j * strideX + f);
#pragma unroll
    for (int i = 0; i < TI; ++i) z[i] = ld4(Za + i * strideZ + f);
#pragma unroll
    for (int j = 0; j < TO; ++j)
#pragma unroll
      for (int i = 0; i < TI; ++i) {
        acc[j][i] = ffma2(make_float2(x[j].x, x[j].y), make_float2(z[i].x, z[i].y), acc[j][i]);
        acc[j][i] = ffma2(make_float2(x[j].z, x[j].w), make_float2(z[i].z, z[i].w), acc[j][i]);
      }
#pragma unroll
    for (int j = 0; j < TO; ++j) x[j] = ld4(Xb + j * strideX + f);
#pragma unroll
    for (int i = 0; i < TI; ++i) z[i] = ld4(Zb + i * strideZ + f);
#pragma unroll
    for (int j = 0; j < TO; ++j)
#pragma unroll
      for (int i = 0; i < TI; ++i) {
        acc[j][i] = ffma2(make_float2(x[j].x, x[j].y), make_float2(z[i].x, z[i].y), acc[j][i]);
        acc[j][i] = ffma2(make_float2(x[j].z, x[j].w), make_float2(z[i].z, z[i].w), acc[j][i]);
      }
  }
}

// MODE 0: 4x12 tile, 30 lanes (pass 2b as built).  MODE 1: 2x12 tile on 2x the lanes-groups (all 32 lanes busy: 10 og x 3 ig... no:
// 5 og(4 rows) -> here 10 og (2 rows) x 3 ig (24 cols)).  MODE 2: 4x6 tile, two column groups per lane sequentially.
template <int MODE, int UNROLL, int T>
__global__ void __launch_bounds__(T, 1) probe(float* sink, int iters) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int rows = 72 + 72 + 40;
  float* Rr = sm + warp * rows * RP;
  float* Vr = Rr + 72 * RP;
  float* Xr = Vr + 72 * RP;
  for (int i = threadIdx.x; i < (T / 32) * rows * RP; i += T) sm[i] = 1e-3f * (i % 97);
  __syncthreads();
  float s = 0.f;
  if (MODE == 0) {
    const int og = lane / 6, ig = lane % 6;
    float2 acc[4][12];
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 12; ++i) acc[j][i] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
      if (og < 5) outer_tile<4, 12, UNROLL>(acc, Xr + og * RP, Rr + ig * RP, Xr + (20 + og) * RP, Vr + ig * RP, 5 * RP, 6 * RP);
      __syncwarp();
    }
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 12; ++i) s += acc[j][i].x + acc[j][i].y;
  } else {
    // 4 x 6 tiles, lane -> (og 0..4, ig 0..5), two passes over the column halves (rows ig + 6 i, i < 6 and i >= 6)
    const int og = lane / 6, ig = lane % 6;
    float2 acc[2][4][6];
    for (int h = 0; h < 2; ++h) for (int j = 0; j < 4; ++j) for (int i = 0; i < 6; ++i) acc[h][j][i] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
      if (og < 5) {
        outer_tile<4, 6, UNROLL>(acc[0], Xr + og * RP, Rr + ig * RP, Xr + (20 + og) * RP, Vr + ig * RP, 5 * RP, 6 * RP);
        outer_tile<4, 6, UNROLL>(acc[1], Xr + og * RP, Rr + (36 + ig) * RP, Xr + (20 + og) * RP, Vr + (36 + ig) * RP, 5 * RP, 6 * RP);
      }
      __syncwarp();
    }
    for (int h = 0; h < 2; ++h) for (int j = 0; j < 4; ++j) for (int i = 0; i < 6; ++i) s += acc[h][j][i].x + acc[h][j][i].y;
  }
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
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;
  float* sink;
  cudaMalloc(&sink, 64);
  const int sms = p.multiProcessorCount, it = 2000;
#define RUN(MODE, UNROLL, T)                                                                                   \
  {                                                                                                            \
    const size_t smem = (size_t)(T / 32) * 184 * RP * 4;                                                       \
    cudaFuncSetAttribute(probe<MODE, UNROLL, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    float ms = time_ms([&] { probe<MODE, UNROLL, T><<<sms, T, smem>>>(sink, it); });                           \
    /* useful MACs per warp per iteration: 20 x 66 x 32 frames x 2 operand pairs */                           \
    printf("mode %d unroll %d threads %3d: %6.1f useful FMA/clk/SM (%s)\n", MODE, UNROLL, T,                   \
           20.0 * 66 * 32 * 2 * it * (T / 32) / (ms * 1e-3 * clk), cudaGetErrorString(cudaGetLastError()));   \
  }
  RUN(0, 1, 128) RUN(0, 1, 256) RUN(0, 2, 128) RUN(0, 2, 256) RUN(1, 1, 128) RUN(1, 1, 256) RUN(1, 2, 256)
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
