// mn_major_probe.cu -- which shared-memory descriptor makes tcgen05.mma.kind::tf32 read an MN-major, 128-byte-swizzled operand
// (rows of 128 bytes = 32 consecutive m (n) values, consecutive rows = consecutive k)?  One CTA, one 128 x 128 x 32 block
// (4 MMAs of K = 8), operands filled with small integers (exact in TF32), result compared with the CPU for every combination of
//   which operand is MN-major (A, B, both)  x  (leading byte offset, stride byte offset) candidates.
// The MN-major operand is stored as four 4 KB pieces (32 m values each): piece j at j * 4096, inside a piece row k at
// (k >> 3) * 1024 + (k & 7) * 128, 16-byte chunk c of the row at (c ^ (k & 7)) << 4 -- i.e. exactly a K-major image tile of the
// TRANSPOSED operand, which is how the training step wants to reuse its activation images.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o profiles/_build/mn_major_probe profiles/micro/mn_major_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 24); ++spin) {
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
__host__ __device__ inline uint32_t sw_off(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

struct Variant {
  int a_mn, b_mn;
  uint32_t lbo, sbo, kstep, layout;   // of an MN-major operand
  uint32_t k_lbo, k_sbo, k_kstep, k_layout;   // of a K-major operand
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ Aimg, const float* __restrict__ Bimg, Variant v,
                                                       float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_slot;
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 4096; i += 128) {
    reinterpret_cast<float*>(tiles)[i] = Aimg[i];
    reinterpret_cast<float*>(tiles + 16384)[i] = Bimg[i];
  }
  if (tid == 0) {
    mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) |
                           (v.a_mn ? 1u << 15 : 0u) | (v.b_mn ? 1u << 16 : 0u);
    const uint32_t a0 = smem_u32(tiles), b0 = smem_u32(tiles + 16384);
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t da = v.a_mn ? make_desc(a0 + kk * v.kstep, v.lbo, v.sbo, v.layout) : make_desc(a0 + kk * v.k_kstep, v.k_lbo, v.k_sbo, v.k_layout);
      const uint64_t db = v.b_mn ? make_desc(b0 + kk * v.kstep, v.lbo, v.sbo, v.layout) : make_desc(b0 + kk * v.k_kstep, v.k_lbo, v.k_sbo, v.k_layout);
      const uint32_t acc = kk != 0;
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "setp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
          "}\n" ::"r"(tmem_d),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
  }
  mbar_wait(&done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 128; c0 += 16) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tmem_d + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) D[(size_t)tid * 128 + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(128u) : "memory");
}

int main() {
  // logical operands: A[m][k], B[n][k], m, n < 128, k < 32
  std::vector<float> A(128 * 32), B(128 * 32);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 32; ++k) A[m * 32 + k] = (float)((m * 7 + k * 3) % 11 - 5), B[m * 32 + k] = (float)((m * 5 + k * 13) % 9 - 4);
  std::vector<float> ref(128 * 128);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      float s = 0;
      for (int k = 0; k < 32; ++k) s += A[m * 32 + k] * B[n * 32 + k];
      ref[m * 128 + n] = s;
    }
  // arrangement 0: 128-byte swizzle (K-major image tile; MN-major = the K-major tile of the transposed operand, 4 KB pieces)
  // arrangement 1: no swizzle, 8 x 16-byte core matrices; K-major: cores along k 128 B apart, 8-row groups 1024 B apart;
  //                MN-major: cores along m 128 B apart, 8-k groups 4096 B apart
  // arrangement 2: MN-major only, 128-byte rows with the 32-byte-atom swizzle: row k at k * 128, 32-byte chunk c at (c ^ (k & 3)) << 5
  auto image = [&](const std::vector<float>& X, int mn, int arr) {
    std::vector<float> img(4096);
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < 32; ++k) {
        uint32_t off;
        if (arr == 0) {
          if (!mn) off = sw_off(r, k >> 2) + 4 * (k & 3);
          else off = (r >> 5) * 4096 + sw_off(k, (r & 31) >> 2) + 4 * (r & 3);
        } else if (arr == 1) {
          if (!mn) off = (r >> 3) * 1024 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4;
          else off = (r >> 2) * 128 + (k >> 3) * 4096 + (k & 7) * 16 + (r & 3) * 4;
        } else {
          const int c32 = (r & 31) >> 3;
          off = (r >> 5) * 4096 + k * 128 + ((c32 ^ (k & 3)) << 5) + 4 * (r & 7);
        }
        img[off / 4] = X[r * 32 + k];
      }
    return img;
  };
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, 16384));
  CK(cudaMalloc(&dB, 16384));
  CK(cudaMalloc(&dD, 128 * 128 * 4));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16384 + 1024));
  struct Case {
    const char* name;
    int a_mn, b_mn, arr;
    uint32_t lbo, sbo, kstep, layout, k_lbo, k_sbo, k_kstep, k_layout;
  };
  const Case cases[] = {
      {"K/K swizzle128", 0, 0, 0, 0, 0, 0, 0, 16, 1024, 32, 2},
      {"K/K interleave", 0, 0, 1, 0, 0, 0, 0, 128, 1024, 256, 0},
      {"MN/K interleave", 1, 0, 1, 4096, 128, 4096, 0, 128, 1024, 256, 0},
      {"K/MN interleave", 0, 1, 1, 4096, 128, 4096, 0, 128, 1024, 256, 0},
      {"MN/MN interleave", 1, 1, 1, 4096, 128, 4096, 0, 128, 1024, 256, 0},
      {"MN/K interleave lbo<->sbo", 1, 0, 1, 128, 4096, 4096, 0, 128, 1024, 256, 0},
      {"MN/K sw128-32B lbo4096 sbo512", 1, 0, 2, 4096, 512, 1024, 1, 16, 1024, 32, 2},
      {"MN/K sw128-32B lbo512 sbo4096", 1, 0, 2, 512, 4096, 1024, 1, 16, 1024, 32, 2},
      {"MN/MN sw128-32B lbo4096 sbo512", 1, 1, 2, 4096, 512, 1024, 1, 16, 1024, 32, 2},
      {"MN/K sw128-32B lbo4096 sbo1024", 1, 0, 2, 4096, 1024, 1024, 1, 16, 1024, 32, 2},
  };
  for (const Case& c : cases) {
    Variant v{c.a_mn, c.b_mn, c.lbo, c.sbo, c.kstep, c.layout, c.k_lbo, c.k_sbo, c.k_kstep, c.k_layout};
    const int k_arr = c.arr == 2 ? 0 : c.arr;   // K-major operands of the 32-byte-atom cases use the ordinary swizzled tile
    std::vector<float> ia = image(A, c.a_mn, c.a_mn ? c.arr : k_arr), ib = image(B, c.b_mn, c.b_mn ? c.arr : k_arr), got(128 * 128);
    CK(cudaMemcpy(dA, ia.data(), 16384, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, ib.data(), 16384, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, 128 * 128 * 4));
    probe_kernel<<<1, 128, 2 * 16384 + 1024>>>(dA, dB, v, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("{\"case\": \"%s\", \"error\": \"%s\"}\n", c.name, cudaGetErrorString(e));
      return 1;
    }
    CK(cudaMemcpy(got.data(), dD, 128 * 128 * 4, cudaMemcpyDeviceToHost));
    int bad = 0, zeros = 0;
    for (int i = 0; i < 128 * 128; ++i) bad += got[i] != ref[i], zeros += got[i] == 0.0f;
    printf("{\"case\": \"%s\", \"mismatches\": %d, \"zeros\": %d, \"d00\": %g, \"ref00\": %g}\n", c.name, bad, zeros, got[0], ref[0]);
  }
  return 0;
}
