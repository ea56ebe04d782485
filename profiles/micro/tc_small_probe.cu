// tc_small_probe.cu -- go / no-go measurement: can the 20-wide layers of the eigenfunction networks run on the
// 5th-generation tensor cores at fp32 parity faster than the FFMA2 thread-private path?
//
// Work item: the forward chain of ONE network on a tile of 128 frames, [128 x 72] -> tanh [128 x 20] -> tanh [128 x 20] -> tanh
// [128 x 20] (the C3 shape with d_r = 66 padded to 72; widths padded to N = 32, K = 24 where the tensor core needs it).
//   * tensor-core version: every layer is D[128 x 32] = A[128 x K] W^T on tcgen05.mma.kind::tf32 with the 3 x TF32 split that holds
//     fp32 parity (hi*hi + lo*hi + hi*lo, fp32 accumulation in tensor memory); operands K-major with the 128-byte swizzle in
//     shared memory; the epilogue (tcgen05.ld -> + bias -> tanh -> split -> st.shared of the next layer's operand rows) runs on 128
//     threads per tile, one frame per thread; two tiles in flight per CTA (8 epilogue warps + 1 MMA-issuing warp) so that one tile's
//     epilogue overlaps the other's MMAs.
//   * SIMT version: the inner loops of cvf_eigen_fast.cu's pass 1 (two frames per thread, weights broadcast from shared memory,
//     packed FFMA2), same tanh.
// Both run persistent CTAs on shared-memory-resident inputs (no HBM traffic): this is the compute ceiling of each formulation.
// Output: frames/s of "one network forward", fp32-equivalent TFLOP/s counting the 2 240 useful FMAs per frame, and the ratio.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o profiles/_build/tc_small_probe profiles/micro/tc_small_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int D0 = 72, H = 20, NP = 32;   // input width (padded), hidden width, padded N of every MMA
constexpr int KB0 = 3;                    // k-blocks (32 floats) of the first layer's operand
constexpr int TILE = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 26); ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major operand tile, 128-byte swizzle: 8-row groups of 1024 B, 16-byte chunk index XOR (row mod 8)
__device__ __host__ __forceinline__ uint32_t sw_off(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// D fp32, A and B TF32, both K-major, N = 32, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}
// the step kernels' tanh (csrc/cvf_common.cuh)
__device__ __forceinline__ float cvf_tanh(float x) {
  const float ax = fabsf(x), s = x * x;
  float p = fmaf(-0.00622109929f, s, 0.0210381374f);
  p = fmaf(p, s, -0.0538453273f);
  p = fmaf(p, s, 0.133325338f);
  p = fmaf(p, s, -0.333333164f);
  const float small = fmaf(x * s, p, x);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.885390082f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  const float big = copysignf(fmaf(-2.0f, r, 1.0f), x);
  return ax < 0.55f ? small : big;
}

// parameters in global memory: W1 [H][D0], b1 [H], W2 [H][H], b2, W3 [H][H], b3 ; input X [TILE][D0]
struct Params {
  const float *W1, *b1, *W2, *b2, *W3, *b3, *X;
  float* out;   // [grid][2][TILE][H] last activations of the last tile of each slot
  int iters;    // tiles per slot per CTA
};

// ------------------------------------------------------------------------------------------------ tensor-core version
// shared memory (bytes), all operand tiles 1024-byte aligned
constexpr int kA0 = 0;                                  // A0 hi | lo: KB0 tiles of [128][32] each
constexpr int kA0Bytes = 2 * KB0 * TILE * 128;          // 98 304
constexpr int kW1 = kA0 + kA0Bytes;                     // W1 hi | lo: KB0 tiles of [32][32]
constexpr int kW1Bytes = 2 * KB0 * NP * 128;            // 24 576
constexpr int kW23 = kW1 + kW1Bytes;                    // W2 hi, W2 lo, W3 hi, W3 lo: [32][32] each
constexpr int kW23Bytes = 4 * NP * 128;                 // 16 384
constexpr int kAct = kW23 + kW23Bytes;                  // per slot: A hi | lo [128][32]
constexpr int kActBytes = 2 * TILE * 128;               // 32 768 per slot
constexpr int kSmemTc = kAct + 2 * kActBytes + 1024;

__global__ void __launch_bounds__(288, 1) tc_forward_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t op_ready[2];    // 128 epilogue threads of a slot have written the next operand
  __shared__ __align__(8) uint64_t acc_ready[2];   // the MMAs of a slot's layer have landed in tensor memory
  __shared__ uint32_t tmem_slot;
  __shared__ float bias[3][NP];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- one-time staging: zero everything, then the split operands
  for (int i = tid; i < (kSmemTc - 1024) / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.0f;
  if (tid < 3 * NP) bias[tid / NP][tid % NP] = 0.0f;
  __syncthreads();
  for (int i = tid; i < TILE * D0; i += blockDim.x) {
    const int row = i / D0, k = i - row * D0;
    float hi, lo;
    split_tf32(p.X[i], hi, lo);
    const uint32_t o = (uint32_t)(k >> 5) * (TILE * 128) + sw_off(row, (k & 31) >> 2) + 4 * (k & 3);
    *reinterpret_cast<float*>(sm + kA0 + o) = hi;
    *reinterpret_cast<float*>(sm + kA0 + KB0 * TILE * 128 + o) = lo;
  }
  for (int i = tid; i < H * D0; i += blockDim.x) {
    const int row = i / D0, k = i - row * D0;
    float hi, lo;
    split_tf32(p.W1[i], hi, lo);
    const uint32_t o = (uint32_t)(k >> 5) * (NP * 128) + sw_off(row, (k & 31) >> 2) + 4 * (k & 3);
    *reinterpret_cast<float*>(sm + kW1 + o) = hi;
    *reinterpret_cast<float*>(sm + kW1 + KB0 * NP * 128 + o) = lo;
  }
  for (int i = tid; i < 2 * H * H; i += blockDim.x) {
    const int l = i / (H * H), r = i - l * H * H, row = r / H, k = r - row * H;
    float hi, lo;
    split_tf32((l == 0 ? p.W2 : p.W3)[r], hi, lo);
    const uint32_t o = sw_off(row, k >> 2) + 4 * (k & 3);
    *reinterpret_cast<float*>(sm + kW23 + (2 * l) * NP * 128 + o) = hi;
    *reinterpret_cast<float*>(sm + kW23 + (2 * l + 1) * NP * 128 + o) = lo;
  }
  if (tid < H) bias[0][tid] = p.b1[tid], bias[1][tid] = p.b2[tid], bias[2][tid] = p.b3[tid];
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) mbar_init(&op_ready[s], 128), mbar_init(&acc_ready[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;

  if (warp == 8) {
    // ---- MMA issuer: for every tile, layer and slot, wait for the operand, issue the split products, commit
    uint32_t ph[2] = {0, 0};
    for (int it = 0; it < p.iters; ++it) {
      for (int layer = 0; layer < 3; ++layer) {
        for (int s = 0; s < 2; ++s) {
          if (layer > 0) {   // operand of layers 2, 3: written by the slot's epilogue threads
            mbar_wait(&op_ready[s], ph[s]);
            ph[s] ^= 1;
          } else if (it > 0) {   // the slot's accumulator was read by the last epilogue of the previous tile
            mbar_wait(&op_ready[s], ph[s]);
            ph[s] ^= 1;
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (lane == 0) {
            const uint32_t acc = tmem_d + 32u * s;
            if (layer == 0) {
#pragma unroll
              for (int kb = 0; kb < KB0; ++kb) {
                const uint64_t a_hi = umma_desc(smem_u32(sm + kA0 + kb * TILE * 128)), a_lo = umma_desc(smem_u32(sm + kA0 + (KB0 + kb) * TILE * 128));
                const uint64_t b_hi = umma_desc(smem_u32(sm + kW1 + kb * NP * 128)), b_lo = umma_desc(smem_u32(sm + kW1 + (KB0 + kb) * NP * 128));
                const int nk = kb == KB0 - 1 ? (D0 - 32 * kb) / 8 : 4;
                for (int kk = 0; kk < nk; ++kk) {
                  const uint64_t ko = (uint64_t)(kk * 32 >> 4);
                  umma_tf32(acc, a_hi + ko, b_hi + ko, (kb | kk) != 0);
                  umma_tf32(acc, a_lo + ko, b_hi + ko, 1);
                  umma_tf32(acc, a_hi + ko, b_lo + ko, 1);
                }
              }
            } else {
              const uint64_t a_hi = umma_desc(smem_u32(sm + kAct + s * kActBytes)), a_lo = umma_desc(smem_u32(sm + kAct + s * kActBytes + TILE * 128));
              const uint64_t b_hi = umma_desc(smem_u32(sm + kW23 + (2 * (layer - 1)) * NP * 128));
              const uint64_t b_lo = umma_desc(smem_u32(sm + kW23 + (2 * (layer - 1) + 1) * NP * 128));
#pragma unroll
              for (int kk = 0; kk < 3; ++kk) {   // K = 24 of the 32 columns
                const uint64_t ko = (uint64_t)(kk * 32 >> 4);
                umma_tf32(acc, a_hi + ko, b_hi + ko, kk != 0);
                umma_tf32(acc, a_lo + ko, b_hi + ko, 1);
                umma_tf32(acc, a_hi + ko, b_lo + ko, 1);
              }
            }
            umma_commit(&acc_ready[s]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue threads: slot = warp / 4, row = 32 (warp % 4) + lane
    const int s = warp >> 2, row = 32 * (warp & 3) + lane;
    uint8_t* act = sm + kAct + s * kActBytes;
    uint32_t ph = 0;
    float last[H];
    for (int it = 0; it < p.iters; ++it) {
      for (int layer = 0; layer < 3; ++layer) {
        mbar_wait(&acc_ready[s], ph);
        ph ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + 32u * s;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float a[H];
#pragma unroll
        for (int c = 0; c < H; ++c) a[c] = cvf_tanh(__uint_as_float(v[c]) + bias[layer][c]);
        if (layer < 2) {
#pragma unroll
          for (int c4 = 0; c4 < H / 4; ++c4) {
            float h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_tf32(a[4 * c4 + e], h[e], l[e]);
            const uint32_t o = sw_off(row, c4);
            *reinterpret_cast<float4*>(act + o) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(act + TILE * 128 + o) = make_float4(l[0], l[1], l[2], l[3]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        } else {
#pragma unroll
          for (int c = 0; c < H; ++c) last[c] = a[c];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (layer < 2 || it + 1 < p.iters) mbar_arrive(&op_ready[s]);
      }
    }
    float* o = p.out + ((size_t)(blockIdx.x * 2 + s) * TILE + row) * H;
#pragma unroll
    for (int c = 0; c < H; ++c) o[c] = last[c];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(64u) : "memory");
}

// ------------------------------------------------------------------------------------------------ SIMT version
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
// 256 threads, two frames per thread (a 512-frame tile = four 128-frame work items), weights k-major in shared memory
__global__ void __launch_bounds__(256, 1) simt_forward_kernel(const Params p) {
  extern __shared__ __align__(16) float smf[];
  float* W1T = smf;                 // [D0][H]
  float* W2T = W1T + D0 * H;        // [H][H]
  float* W3T = W2T + H * H;
  float* bs = W3T + H * H;          // [3][H]
  float* tile = bs + 3 * H + 4;     // [D0][512]
  const int tid = threadIdx.x;
  for (int i = tid; i < D0 * H; i += 256) W1T[i] = p.W1[(i % H) * D0 + i / H];
  for (int i = tid; i < H * H; i += 256) W2T[i] = p.W2[(i % H) * H + i / H], W3T[i] = p.W3[(i % H) * H + i / H];
  if (tid < H) bs[tid] = p.b1[tid], bs[H + tid] = p.b2[tid], bs[2 * H + tid] = p.b3[tid];
  for (int i = tid; i < D0 * 512; i += 256) tile[i] = p.X[((i % 512) % TILE) * D0 + i / 512];
  __syncthreads();
  const int c0 = 2 * tid;
  float a[2][H];
  for (int it = 0; it < p.iters; ++it) {
    float2 z[2][H / 2];
#pragma unroll
    for (int j = 0; j < H / 2; ++j) z[0][j] = z[1][j] = *reinterpret_cast<const float2*>(bs + 2 * j);
#pragma unroll 2
    for (int kk = 0; kk < D0; ++kk) {
      const float2 x = *reinterpret_cast<const float2*>(tile + kk * 512 + c0);
      const float2 x0 = make_float2(x.x, x.x), x1 = make_float2(x.y, x.y);
#pragma unroll
      for (int q = 0; q < H / 4; ++q) {
        const float4 wv = *reinterpret_cast<const float4*>(W1T + kk * H + 4 * q);
        z[0][2 * q] = ffma2(make_float2(wv.x, wv.y), x0, z[0][2 * q]);
        z[0][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), x0, z[0][2 * q + 1]);
        z[1][2 * q] = ffma2(make_float2(wv.x, wv.y), x1, z[1][2 * q]);
        z[1][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), x1, z[1][2 * q + 1]);
      }
    }
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int j = 0; j < H / 2; ++j) a[f][2 * j] = cvf_tanh(z[f][j].x), a[f][2 * j + 1] = cvf_tanh(z[f][j].y);
#pragma unroll
    for (int l = 1; l < 3; ++l) {
      const float* wt = l == 1 ? W2T : W3T;
#pragma unroll
      for (int j = 0; j < H / 2; ++j) z[0][j] = z[1][j] = *reinterpret_cast<const float2*>(bs + l * H + 2 * j);
#pragma unroll
      for (int kk = 0; kk < H; ++kk) {
        const float2 x0 = make_float2(a[0][kk], a[0][kk]), x1 = make_float2(a[1][kk], a[1][kk]);
#pragma unroll
        for (int q = 0; q < H / 4; ++q) {
          const float4 wv = *reinterpret_cast<const float4*>(wt + kk * H + 4 * q);
          z[0][2 * q] = ffma2(make_float2(wv.x, wv.y), x0, z[0][2 * q]);
          z[0][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), x0, z[0][2 * q + 1]);
          z[1][2 * q] = ffma2(make_float2(wv.x, wv.y), x1, z[1][2 * q]);
          z[1][2 * q + 1] = ffma2(make_float2(wv.z, wv.w), x1, z[1][2 * q + 1]);
        }
      }
#pragma unroll
      for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int j = 0; j < H / 2; ++j) a[f][2 * j] = cvf_tanh(z[f][j].x), a[f][2 * j + 1] = cvf_tanh(z[f][j].y);
    }
    // keep the result alive and perturb the next iteration's input so that nothing is hoisted
    tile[(it % D0) * 512 + c0] += 1e-12f * a[0][0];
  }
  if (tid < 64) {
    float* o = p.out + ((size_t)(blockIdx.x * 2) * TILE + c0) * H;
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
      for (int c = 0; c < H; ++c) o[f * H + c] = a[f][c];
  }
}

// ------------------------------------------------------------------------------------------------ host
static void cpu_forward(const std::vector<float>& X, const std::vector<float>* W, const std::vector<float>* b, int row, double* out) {
  double a0[D0], a1[H], a2[H];
  for (int k = 0; k < D0; ++k) a0[k] = X[row * D0 + k];
  for (int o = 0; o < H; ++o) {
    double z = b[0][o];
    for (int k = 0; k < D0; ++k) z += (double)W[0][o * D0 + k] * a0[k];
    a1[o] = tanh(z);
  }
  for (int o = 0; o < H; ++o) {
    double z = b[1][o];
    for (int k = 0; k < H; ++k) z += (double)W[1][o * H + k] * a1[k];
    a2[o] = tanh(z);
  }
  for (int o = 0; o < H; ++o) {
    double z = b[2][o];
    for (int k = 0; k < H; ++k) z += (double)W[2][o * H + k] * a2[k];
    out[o] = tanh(z);
  }
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  srand(1);
  auto rnd = [] { return (float)rand() / RAND_MAX * 2.0f - 1.0f; };
  std::vector<float> X(TILE * D0), W[3], b[3];
  for (auto& v : X) v = 3.0f * rnd();
  for (int k = 66; k < D0; ++k)
    for (int r = 0; r < TILE; ++r) X[r * D0 + k] = 0.0f;   // padding columns of d_r = 66
  W[0].resize(H * D0), W[1].resize(H * H), W[2].resize(H * H);
  for (auto& v : W[0]) v = rnd() / sqrtf(66.0f);
  for (auto& v : W[1]) v = rnd() / sqrtf((float)H);
  for (auto& v : W[2]) v = rnd() / sqrtf((float)H);
  for (int l = 0; l < 3; ++l) {
    b[l].resize(H);
    for (auto& v : b[l]) v = 0.2f * rnd();
  }
  float *dX, *dW[3], *db[3], *dout;
  CK(cudaMalloc(&dX, X.size() * 4));
  CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  for (int l = 0; l < 3; ++l) {
    CK(cudaMalloc(&dW[l], W[l].size() * 4));
    CK(cudaMemcpy(dW[l], W[l].data(), W[l].size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&db[l], H * 4));
    CK(cudaMemcpy(db[l], b[l].data(), H * 4, cudaMemcpyHostToDevice));
  }
  const size_t out_floats = (size_t)sms * 2 * TILE * H;
  CK(cudaMalloc(&dout, out_floats * 4));
  Params p{dW[0], db[0], dW[1], db[1], dW[2], db[2], dX, dout, iters};
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<float> out(out_floats);
  const double fma_per_frame = 66.0 * H + 2.0 * H * H;   // useful FMAs of one network forward

  // ---- tensor cores
  CK(cudaFuncSetAttribute(tc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTc));
  Params warm = p;
  warm.iters = 10;
  tc_forward_kernel<<<sms, 288, kSmemTc>>>(warm);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  tc_forward_kernel<<<sms, 288, kSmemTc>>>(p);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms_tc = 0;
  CK(cudaEventElapsedTime(&ms_tc, e0, e1));
  CK(cudaMemcpy(out.data(), dout, out_floats * 4, cudaMemcpyDeviceToHost));
  double err_tc = 0;
  for (int row = 0; row < TILE; ++row) {
    double ref[H];
    cpu_forward(X, W, b, row, ref);
    for (int c = 0; c < H; ++c)
      for (int s = 0; s < 2; ++s) err_tc = fmax(err_tc, fabs(out[((size_t)s * TILE + row) * H + c] - ref[c]));
  }
  const double frames_tc = (double)sms * 2 * TILE * iters;
  // ---- SIMT
  const size_t smem_simt = (size_t)(D0 * H + 2 * H * H + 3 * H + 4 + D0 * 512) * 4;
  CK(cudaFuncSetAttribute(simt_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_simt));
  simt_forward_kernel<<<sms, 256, smem_simt>>>(warm);
  CK(cudaDeviceSynchronize());
  Params ps = p;
  ps.iters = iters / 2;   // a SIMT iteration is 512 frames, a tensor-core iteration 256
  CK(cudaEventRecord(e0));
  simt_forward_kernel<<<sms, 256, smem_simt>>>(ps);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms_simt = 0;
  CK(cudaEventElapsedTime(&ms_simt, e0, e1));
  CK(cudaMemcpy(out.data(), dout, out_floats * 4, cudaMemcpyDeviceToHost));
  double err_simt = 0;
  for (int row = 0; row < 128; ++row) {
    double ref[H];
    cpu_forward(X, W, b, row % TILE, ref);
    for (int c = 0; c < H; ++c) err_simt = fmax(err_simt, fabs(out[(size_t)row * H + c] - ref[c]));
  }
  const double frames_simt = (double)sms * 512 * ps.iters;
  const double tf_tc = frames_tc * fma_per_frame * 2 / (ms_tc * 1e-3) / 1e12, tf_simt = frames_simt * fma_per_frame * 2 / (ms_simt * 1e-3) / 1e12;
  printf("{\"probe\": \"tc_small_probe\", \"sms\": %d, \"clock_mhz\": %.0f, \"shape\": \"[128 x 72] -> 20 -> 20 -> 20, tanh, one network forward\",\n", sms,
         khz / 1e3);
  printf(" \"tensor_core_3xtf32\": {\"ms\": %.3f, \"frames_per_s\": %.4g, \"fp32_equiv_tflops\": %.2f, \"fma_per_clk_per_sm\": %.1f, \"max_abs_err\": %.2e},\n",
         ms_tc, frames_tc / (ms_tc * 1e-3), tf_tc, frames_tc * fma_per_frame / (ms_tc * 1e-3) / sms / (khz * 1e3), err_tc);
  printf(" \"simt_ffma2\": {\"ms\": %.3f, \"frames_per_s\": %.4g, \"fp32_equiv_tflops\": %.2f, \"fma_per_clk_per_sm\": %.1f, \"max_abs_err\": %.2e},\n", ms_simt,
         frames_simt / (ms_simt * 1e-3), tf_simt, frames_simt * fma_per_frame / (ms_simt * 1e-3) / sms / (khz * 1e3), err_simt);
  printf(" \"ratio_tc_over_simt\": %.2f}\n", tf_tc / tf_simt);
  return 0;
}
