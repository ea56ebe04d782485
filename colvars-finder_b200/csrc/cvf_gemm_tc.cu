// cvf_gemm_tc.cu -- C[M x N] = Aop[M x K] Bop[K x N] in fp32 accuracy on the 5th-generation tensor cores (tcgen05).
//
// fp32 operands are split, x = hi + lo with hi = x truncated to TF32 (10-bit mantissa), and every 128 x 128 x 32 block is three
// kind::tf32 products accumulated in fp32 in tensor memory:
//     D += Ahi Bhi + Alo Bhi + Ahi Blo          (the dropped lo*lo term is 2^-20 relative)
// which keeps the 1e-5 parity target of the training step that a single TF32 pass (2^-11) cannot.
//
//   * operands: both operands arrive as "tile images" (cvf_gemm.cuh): already split into hi / lo and laid out K-major ([row][32 k] =
//     128-byte rows) in the canonical 128-byte-swizzled layout the UMMA shared-memory descriptor expects (8-row groups of 1024
//     bytes, 16-byte chunk index XOR row mod 8), so a stage (Ahi | Alo | Bhi | Blo, 64 KB) is filled by two 1-D bulk copies (TMA)
//     issued by one thread; weights' images are built once per step (tile_image_kernel), activations' and deltas' images are
//     written by the epilogue of the product that computes them;
//   * persistent, warp-specialised (see the kernel): a copy thread, an MMA thread (M = 128, N = 128, K = 8 per instruction; 12 per
//     stage; three stages handed over through mbarriers) and eight drain / epilogue warps; the epilogue of a tile overlaps the
//     first two chains of the next tile;
//   * the 128 x 128 fp32 accumulator lives in tensor memory and is read back with tcgen05.ld (32 lanes x 32 columns per warp
//     instruction).  The tensor core adds into it with truncation, not round-to-nearest: the error is a bias that grows with
//     the number of accumulated MMAs (measured: 3e-5 relative after the 1128 MMAs of a K = 3000 product, which broke the 2e-5
//     gradient bar at the real C5 shape).  So a chain is at most kGroup k-blocks (96 MMAs) long: two chain buffers (2 x 128
//     columns) alternate, and while the tensor core fills one the drain warps add the other, 16 columns at a time and with
//     ordinary round-to-nearest additions in registers, onto a third 128-column "master" tile that also lives in tensor
//     memory (register pressure stays what it was); the fused epilogue (bias / tanh / multiplication by 1 - act^2) reads the
//     master tile.
#include <stdint.h>

#include "cvf_common.cuh"
#include "cvf_gemm.cuh"

namespace cvf {
namespace wide {

constexpr int TM = 128, TN = 128, TK = 32;            // CTA tile; TK fp32 = one 128-byte swizzle row
constexpr int kTileBytes = TM * TK * 4;               // 16 KB per operand tile
constexpr int kStageBytes = 4 * kTileBytes;           // Ahi, Alo, Bhi, Blo
constexpr int kStages = 3;
constexpr int kGroup = 8;                             // k-blocks accumulated in tensor memory before a drain (8 x 12 MMAs, K = 256)
constexpr uint32_t kTmemCols = 512;                   // two chain buffers and the master tile, 128 columns each (power of two: 512)
constexpr uint32_t kMasterCol = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// bounded wait: a descriptor mistake must not hang the GPU
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

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle: start address, stride between 8-row groups 1024 B,
// descriptor version 1 (sm_100), layout type SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // bits 0-13
  d |= (uint64_t)1 << 16;                             // leading byte offset: unused for a swizzled K-major atom (CUTLASS writes 1)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset
  d |= (uint64_t)1 << 46;                             // version
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

// instruction descriptor: D fp32, A and B TF32, both K-major, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

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

// byte offset of (row, 16-byte chunk) inside a K-major 128-byte-swizzled tile
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = x - hi;
}

// ---------------------------------------------------------------------------------------------------- the product kernel
// Persistent: one CTA per SM walks over the output tiles (column tile fastest, so that the CTAs running at one time share A row
// tiles in L2).  Ten warps:
//   warp 9   one thread issues the bulk copies (TMA) that fill the stages from the operand images, as far ahead as the ring allows;
//   warp 8   one thread issues the MMAs of every filled stage, commits the stage back to the copy thread, and commits every chain
//            of kGroup k-blocks to the drain warps;
//   warps 0-7  add finished chains onto the master tile (tensor memory) and, after a tile's last chain, run its epilogue -- while
//            the tensor core is already working on the next tile's first two chains (the two chain buffers).
// The stage ring, the chain buffers and their mbarrier phases run on across tiles.
constexpr int kGemmThreads = 320;

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void bar_sync_epilogue() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// shared -> global bulk copy (TMA), one bulk group per copy; wait until the groups' shared-memory reads / everything is done
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(kGemmThreads, 1) tc_gemm_kernel(const Gemm g, int nx, int ny, int n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_done[kStages];   // tensor core has finished reading the stage
  __shared__ __align__(8) uint64_t full[kStages];       // the stage's bulk copies have landed
  __shared__ __align__(8) uint64_t acc_full[2];         // the MMAs of a chain have all landed in chain buffer b
  __shared__ __align__(8) uint64_t acc_free[2];         // the 256 draining threads have read chain buffer b
  __shared__ uint32_t tmem_base_slot;
  __shared__ double red[2][8];
  // 1024-byte aligned operand area (the swizzle pattern is a function of the absolute address bits)
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&mma_done[s], 1), mbar_init(&full[s], 1);
    for (int b = 0; b < 2; ++b) mbar_init(&acc_full[b], 1), mbar_init(&acc_free[b], 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;
  constexpr size_t kImgTile = 2 * kTileBytes / 4;   // floats of one (row tile, k-block) of an image: hi | lo

  // tile t of this product: column tile x (fastest), row tile y, k-split z; its k-blocks
  auto tile_blocks = [&](int t, int& x, int& y, int& z, int& kbeg) {
    x = t % nx;
    const int r = t / nx;
    y = r % ny, z = r / ny;
    kbeg = z * g.k_per_split;
    const int kend = min(g.K, kbeg + g.k_per_split);
    return (kend - kbeg + TK - 1) / TK;
  };

  if (warp == 9) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int x, y, z, kbeg;
        const int n_blocks = tile_blocks(t, x, y, z, kbeg);
        const float* a_img = g.a_img + ((size_t)y * g.a_img_kblocks + kbeg / TK) * kImgTile;
        const float* b_img = g.b_img + ((size_t)x * g.b_img_kblocks + kbeg / TK) * kImgTile;
        for (int blk = 0; blk < n_blocks; ++blk, ++it) {
          const uint32_t s = it % kStages, use = it / kStages;
          if (use >= 1) mbar_wait(&mma_done[s], (use - 1) & 1);   // the MMAs that last read this stage have finished
          uint8_t* st = tiles + (size_t)s * kStageBytes;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"((uint32_t)(4 * kTileBytes))
                       : "memory");
          bulk_g2s(st, a_img + (size_t)blk * kImgTile, 2 * kTileBytes, &full[s]);
          bulk_g2s(st + 2 * kTileBytes, b_img + (size_t)blk * kImgTile, 2 * kTileBytes, &full[s]);
        }
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      uint32_t it = 0, gi = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int x, y, z, kbeg;
        const int n_blocks = tile_blocks(t, x, y, z, kbeg);
        for (int blk = 0; blk < n_blocks; ++blk, ++it) {
          const uint32_t s = it % kStages, use = it / kStages;
          const int kb = blk % kGroup;
          const uint32_t ab = gi & 1;
          // a chain buffer is reused by chain gi + 2: its previous contents (chain gi - 2) must have been drained
          if (kb == 0 && gi >= 2) mbar_wait(&acc_free[ab], ((gi >> 1) - 1) & 1);
          mbar_wait(&full[s], use & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_acc = tmem_d + 128u * ab;
          const uint32_t a_hi = smem_u32(tiles + (size_t)s * kStageBytes);
          const uint64_t da_hi = umma_desc(a_hi), da_lo = umma_desc(a_hi + kTileBytes), db_hi = umma_desc(a_hi + 2 * kTileBytes),
                         db_lo = umma_desc(a_hi + 3 * kTileBytes);
#pragma unroll
          for (int kk = 0; kk < TK / 8; ++kk) {
            const uint64_t ko = (uint64_t)(kk * 32 >> 4);   // 8 TF32 = 32 bytes along K inside the swizzle row (address field: >> 4)
            umma_tf32(tmem_acc, da_hi + ko, db_hi + ko, (kb | kk) != 0);
            umma_tf32(tmem_acc, da_lo + ko, db_hi + ko, 1);
            umma_tf32(tmem_acc, da_hi + ko, db_lo + ko, 1);
          }
          umma_commit(&mma_done[s]);
          if (kb == kGroup - 1 || blk == n_blocks - 1) {
            umma_commit(&acc_full[ab]);
            ++gi;
          }
        }
      }
    }
  } else {
    // chain gi sits in chain buffer gi & 1; its use number gi >> 1 gives the phase of the buffer's barriers.  A thread owns lanes
    // 32 (warp % 4) .. +31 (rows) x columns 64 (warp / 4) .. +63 of the tile, in the chain buffers and in the master tile alike,
    // so the drains and the epilogue of one thread never touch another thread's part: no barrier between them.
    uint32_t gi = 0;
    const uint32_t lane_base = tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(64 * (warp >> 2));
    uint8_t* slab = tiles + (size_t)kStages * kStageBytes + 4096 * warp;   // this warp's 4 KB for outgoing image pieces
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      int x, y, z, kbeg;
      const int n_blocks = tile_blocks(t, x, y, z, kbeg);
      const int n_groups = (n_blocks + kGroup - 1) / kGroup;
      for (int grp = 0; grp < n_groups; ++grp, ++gi) {
        const uint32_t b = gi & 1;
        mbar_wait(&acc_full[b], (gi >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int piece = 0; piece < 4; ++piece) {
          uint32_t c[16], m[16];
          ld16(lane_base + 128u * b + (uint32_t)(16 * piece), c);
          if (grp > 0) ld16(lane_base + kMasterCol + (uint32_t)(16 * piece), m);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (grp > 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = __float_as_uint(__uint_as_float(c[i]) + __uint_as_float(m[i]));
          }
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
                  lane_base + kMasterCol + (uint32_t)(16 * piece)),
              "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7]), "r"(c[8]), "r"(c[9]), "r"(c[10]),
              "r"(c[11]), "r"(c[12]), "r"(c[13]), "r"(c[14]), "r"(c[15])
              : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&acc_free[b]);
      }

      // ---- epilogue of tile (x, y, z).  Besides C (row-major fp32) it writes the operand images of C that later products read
      // (cvf_gemm.cuh): the thread's 32 values of a row are one 128-byte row of a K-major tile (eight 16-byte chunks at their
      // swizzled places), and for a fixed column the warp's 32 rows are one 128-byte row of the transposed tile (32 scalar stores
      // = one full line).
      const int m0 = y * TM, n0 = x * TN;
      float* C = g.C ? g.C + (size_t)z * g.c_split_stride : nullptr;
      const int row = m0 + 32 * (warp & 3) + lane;
      const bool row_ok = row < g.M;
      float wf = 0.0f, e2 = 0.0f;
      if (g.epi == EPI_BIAS_LOSS && row_ok) wf = __ldg(g.loss_w + row);
#pragma unroll 1
      for (int cb = 0; cb < 2; ++cb) {
        const int col0 = 64 * (warp >> 2) + 32 * cb;
        // everything this half-row needs from memory is requested before the first use, so that one latency is exposed per
        // half-row, not one per element: the accumulators (tensor memory), the bias (the same 32 values for every lane) and the
        // row's 32 values of the side input (activations for 1 - act^2, or the reconstruction target)
        float v[32], bb[32];
        float4 aa[8];
        uint32_t lo[16], hi[16];
        {
          const uint32_t taddr = lane_base + kMasterCol + (uint32_t)(32 * cb);
          ld16(taddr, lo);
          ld16(taddr + 16, hi);
        }
        const bool has_bias = g.epi == EPI_BIAS || g.epi == EPI_BIAS_TANH || g.epi == EPI_BIAS_LOSS;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int n = n0 + col0 + c;
          bb[c] = (has_bias && n < g.N) ? __ldg(g.bias + n) : 0.0f;
        }
        const bool act_from_img = g.epi == EPI_MUL_OM && g.act_img != nullptr;
        const float* side = act_from_img ? nullptr
                            : g.epi == EPI_MUL_OM ? g.act + (size_t)row * g.ldc
                            : g.epi == EPI_BIAS_LOSS ? g.loss_in + (size_t)row * g.loss_ld : nullptr;
        if (act_from_img) {
          // this thread's row of the activations' K-major tile (row tile y, k-block = this half-row's 32 columns): hi + lo
          const int kb = (n0 + col0) >> 5;
          const uint8_t* tile = reinterpret_cast<const uint8_t*>(g.act_img + ((size_t)y * g.act_img_kblocks + kb) * kImgTile);
          const int r = 32 * (warp & 3) + lane;
          const bool in = kb < g.act_img_kblocks;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in) {
              const uint32_t o = sw_off(r, c4);
              const float4 h4 = __ldg(reinterpret_cast<const float4*>(tile + o));
              const float4 l4 = __ldg(reinterpret_cast<const float4*>(tile + kTileBytes + o));
              t4 = make_float4(h4.x + l4.x, h4.y + l4.y, h4.z + l4.z, h4.w + l4.w);
            }
            aa[c4] = t4;
          }
        } else {
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const int n = n0 + col0 + 4 * c4;
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (side != nullptr && row_ok && n < g.N) {
              if (n + 3 < g.N) {
                t4 = __ldg(reinterpret_cast<const float4*>(side + n));
              } else {
                t4.x = __ldg(side + n);
                if (n + 1 < g.N) t4.y = __ldg(side + n + 1);
                if (n + 2 < g.N) t4.z = __ldg(side + n + 2);
              }
            }
            aa[c4] = t4;
          }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(lo[c]), v[16 + c] = __uint_as_float(hi[c]);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const int n = n0 + col0 + 4 * c4;
          float o[4] = {v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]};
          const float sd[4] = {aa[c4].x, aa[c4].y, aa[c4].z, aa[c4].w};
          if (row_ok && n < g.N) {
            if (has_bias) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                o[c] += bb[4 * c4 + c];
                if (g.epi == EPI_BIAS_TANH) o[c] = cvf_tanh(o[c]);
              }
              if (g.epi == EPI_BIAS_LOSS) {   // e = out - in, sum of e^2 per row, delta = 2 w e
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  if (n + c < g.N) {
                    const float e = o[c] - sd[c];
                    e2 = fmaf(e, e, e2);
                    o[c] = 2.0f * wf * e;
                  }
              }
            } else if (g.epi == EPI_MUL_OM) {
#pragma unroll
              for (int c = 0; c < 4; ++c) o[c] *= fmaf(-sd[c], sd[c], 1.0f);
            }
            if (C != nullptr) {
              if (n + 3 < g.N) {
                *reinterpret_cast<float4*>(C + (size_t)row * g.ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
              } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  if (n + c < g.N) C[(size_t)row * g.ldc + n + c] = o[c];
              }
            }
          }
          // what lies outside the matrix is zero in the images
#pragma unroll
          for (int c = 0; c < 4; ++c) v[4 * c4 + c] = (row_ok && n + c < g.N) ? o[c] : 0.0f;
        }
        if (g.c_img_k != nullptr) {
          // K-major image: the warp's 32 rows x 32 columns are a contiguous 4 KB piece (four 8-row groups) of the hi tile and of
          // the lo tile; it is put together in the warp's shared-memory slab (16-byte chunks at their swizzled places,
          // conflict-free) and leaves as one bulk copy (TMA) per piece instead of 32-line scattered stores
          const int kb = (n0 + col0) >> 5;
          if (kb < g.c_img_k_kblocks) {
            uint8_t* tile = reinterpret_cast<uint8_t*>(g.c_img_k + ((size_t)y * g.c_img_k_kblocks + kb) * kImgTile) + 4096 * (warp & 3);
            const uint32_t so = (uint32_t)((lane >> 3) * 1024 + (lane & 7) * 128);
#pragma unroll
            for (int part = 0; part < 2; ++part) {
              if (lane == 0) bulk_wait_read();   // the slab's previous piece has left
              __syncwarp();
#pragma unroll
              for (int ch = 0; ch < 8; ++ch) {
                float h[4], l[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) split_tf32(v[4 * ch + c], h[c], l[c]);
                *reinterpret_cast<float4*>(slab + so + ((ch ^ (lane & 7)) << 4)) =
                    part == 0 ? make_float4(h[0], h[1], h[2], h[3]) : make_float4(l[0], l[1], l[2], l[3]);
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (lane == 0) bulk_s2g(tile + part * kTileBytes, slab, 4096);
            }
          }
        }
        if (g.c_img_t != nullptr) {
          const int kbt = (m0 >> 5) + (warp & 3);
          if (kbt < g.c_img_t_kblocks) {
            uint8_t* tile = reinterpret_cast<uint8_t*>(g.c_img_t + ((size_t)x * g.c_img_t_kblocks + kbt) * kImgTile);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int nl = col0 + c;
              float xv = v[c];
              if (g.c_img_t_ones > 0 && n0 + nl == g.c_img_t_ones && row_ok) xv = 1.0f;
              float h, l;
              split_tf32(xv, h, l);
              const uint32_t o = sw_off(nl, lane >> 2) + 4 * (lane & 3);
              *reinterpret_cast<float*>(tile + o) = h;
              *reinterpret_cast<float*>(tile + kTileBytes + o) = l;
            }
            // the row of ones starts a row tile of its own when N is a multiple of 128: the last column tile's CTAs write it
            if (g.c_img_t_ones == n0 + TN && col0 == 0) {
              uint8_t* t2 = reinterpret_cast<uint8_t*>(g.c_img_t + ((size_t)(x + 1) * g.c_img_t_kblocks + kbt) * kImgTile);
              const uint32_t o = sw_off(0, lane >> 2) + 4 * (lane & 3);
              *reinterpret_cast<float*>(t2 + o) = row_ok ? 1.0f : 0.0f;
              *reinterpret_cast<float*>(t2 + kTileBytes + o) = 0.0f;
            }
          }
        }
      }
      if (g.epi == EPI_BIAS_LOSS) {
        // this tile's share of (sum w |e|^2, sum w): fixed order (lanes by shuffle, then the eight warps), one slot per tile
        double s2 = (double)wf * (double)e2, s0 = (x == 0 && (warp >> 2) == 0) ? (double)wf : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o), s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        if (lane == 0) red[0][warp] = s2, red[1][warp] = s0;
        bar_sync_epilogue();
        if (tid < 2) {
          double tsum = 0.0;
          for (int q = 0; q < 8; ++q) tsum += red[tid][q];
          g.loss_part[(size_t)t * 2 + tid] = tsum;
        }
        bar_sync_epilogue();
      }
    }
  }
  if (warp < 8 && lane == 0) bulk_wait_all();   // outgoing image pieces have left shared memory and are written
  // every warp's tensor-memory reads are done before the allocation is returned
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------- image builder
struct Stage4 {
  float4 v[4];
};

// global -> registers: this thread's 16 floats of a 128 x 32 operand tile (rows = m or n, k0 .. k0+31), zero beyond the edges
__device__ __forceinline__ void tc_load(Stage4& r, const float* __restrict__ P, long long ld, int kcontig, int row0, int nrows, int k0,
                                        int k1, int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kcontig) {   // (row = idx / 8, chunk = idx % 8): 4 consecutive k
      const int idx = tid + 256 * i, row = row0 + (idx >> 3), k = k0 + 4 * (idx & 7);
      if (row < nrows && k < k1) {
        v = __ldg(reinterpret_cast<const float4*>(P + (size_t)row * ld + k));
        if (k + 3 >= k1) {
          if (k + 1 >= k1) v.y = 0.f;
          if (k + 2 >= k1) v.z = 0.f;
          v.w = 0.f;
        }
      }
    } else {         // 4 consecutive rows at one k; a warp covers 16 k x 8 rows (see tc_store)
      const int lane = tid & 31, wq = (tid >> 5) + 8 * (i >> 1);
      const int k = k0 + (lane & 15) + 16 * (i & 1), row = row0 + 4 * ((lane >> 4) + 2 * wq);
      if (k < k1 && row < nrows) {
        v = __ldg(reinterpret_cast<const float4*>(P + (size_t)k * ld + row));
        if (row + 3 >= nrows) {
          if (row + 1 >= nrows) v.y = 0.f;
          if (row + 2 >= nrows) v.z = 0.f;
          v.w = 0.f;
        }
      }
    }
    r.v[i] = v;
  }
}

// registers -> the hi and lo tiles of one operand (K-major, swizzled)
__device__ __forceinline__ void tc_store(uint8_t* hi_tile, uint8_t* lo_tile, const Stage4& r, int kcontig, int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float x[4] = {r.v[i].x, r.v[i].y, r.v[i].z, r.v[i].w};
    float h[4], l[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) split_tf32(x[c], h[c], l[c]);
    if (kcontig) {
      const int idx = tid + 256 * i, row = idx >> 3, chunk = idx & 7;
      const uint32_t o = sw_off(row, chunk);
      *reinterpret_cast<float4*>(hi_tile + o) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(lo_tile + o) = make_float4(l[0], l[1], l[2], l[3]);
    } else {
      // transposing store.  Lanes of a warp differ in k & 15 and in the parity of the row quad, so for each of the four
      // components the 32 scalar stores fall into 32 different banks: word = 4 * ((k >> 2) ^ (row & 7)) + (k & 3) modulo 32
      // takes every value once (k & 3, (k >> 2) & 3 and the row & 4 bit that the XOR moves into bit 2 of the chunk).
      const int lane = tid & 31, wq = (tid >> 5) + 8 * (i >> 1);
      const int k = (lane & 15) + 16 * (i & 1), row = 4 * ((lane >> 4) + 2 * wq);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t o = sw_off(row + c, k >> 2) + 4 * (k & 3);
        *reinterpret_cast<float*>(hi_tile + o) = h[c];
        *reinterpret_cast<float*>(lo_tile + o) = l[c];
      }
    }
  }
}

// One CTA per (row tile, k-block): the same load / split / swizzled store as the product kernel's staging, into shared memory,
// then the finished 32 KB image (hi tile | lo tile) goes out as coalesced 16-byte stores.  ones_row > 0: the operand has one more
// row, all ones (k < K), at that index (>= rows).
__global__ void __launch_bounds__(256) tile_image_kernel(const float* __restrict__ X, long long ld, int kcontig, int rows, int K,
                                                         int kblocks, int row_tiles, int ones_row, float* __restrict__ img) {
  __shared__ __align__(1024) uint8_t buf[2 * kTileBytes];
  const int tid = threadIdx.x;
  const long long n_items = (long long)row_tiles * kblocks;
  for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
    const int rt = (int)(it / kblocks), kb = (int)(it - (long long)rt * kblocks);
    Stage4 r;
    tc_load(r, X, ld, kcontig, rt * TM, rows, kb * TK, K, tid);
    __syncthreads();   // the previous item's copy-out has finished reading buf
    tc_store(buf, buf + kTileBytes, r, kcontig, tid);
    __syncthreads();
    if (ones_row > 0 && ones_row / TM == rt && tid < TK) {   // after the zero padding of the rows beyond `rows` has been stored
      const int k = kb * TK + tid;
      const uint32_t o = sw_off(ones_row - rt * TM, tid >> 2) + 4 * (tid & 3);
      *reinterpret_cast<float*>(buf + o) = k < K ? 1.0f : 0.0f;
    }
    __syncthreads();
    float4* dst = reinterpret_cast<float4*>(img + (size_t)it * (2 * kTileBytes / 4));
    const float4* src = reinterpret_cast<const float4*>(buf);
#pragma unroll
    for (int i = 0; i < 2 * kTileBytes / 16 / 256; ++i) dst[tid + 256 * i] = src[tid + 256 * i];
  }
}

int launch_tile_image(const float* X, long long ld, int kcontig, int rows, int K, int ones_row, float* img, cudaStream_t stream) {
  const int kblocks = (K + TK - 1) / TK;
  const int row_tiles = ((ones_row > 0 ? ones_row + 1 : rows) + TM - 1) / TM;
  const long long n_items = (long long)row_tiles * kblocks;
  const int grid = (int)(n_items > 148 * 16 ? 148 * 16 : n_items);
  CVF_LAUNCH(K_AE_STEP, stream, tile_image_kernel<<<grid, 256, 0, stream>>>(X, ld, kcontig, rows, K, kblocks, row_tiles, ones_row, img));
  CVF_CUDA(cudaGetLastError());
  return 0;
}

int launch_gemm_tc(const Gemm& g, int splits, cudaStream_t stream) {
  if (g.a_img == nullptr || g.b_img == nullptr) {
    set_error("tc_gemm_kernel: both operands must be given as tile images");
    return CVF_E_ARG;
  }
  if (splits > 1 && (g.c_img_k != nullptr || g.c_img_t != nullptr || g.epi == EPI_BIAS_LOSS)) {
    set_error("tc_gemm_kernel: image outputs and the loss epilogue need a single k-split");
    return CVF_E_ARG;
  }
  const size_t smem = (size_t)kStages * kStageBytes + 8 * 4096 + 1024;   // stages, the epilogue warps' slabs, alignment slack
  static bool configured = false;
  if (!configured) {
    CVF_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int nx = (g.N + TN - 1) / TN, ny = (g.M + TM - 1) / TM, n_tiles = nx * ny * splits;
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  CVF_LAUNCH(K_AE_STEP, stream, tc_gemm_kernel<<<grid, kGemmThreads, smem, stream>>>(g, nx, ny, n_tiles));
  {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("tc_gemm_kernel launch failed: grid %d, %zu B shared memory, M %d N %d K %d: %s", grid, smem, g.M, g.N, g.K,
                cudaGetErrorString(e));
      return (int)e;
    }
  }
  return 0;
}

}  // namespace wide
}  // namespace cvf
