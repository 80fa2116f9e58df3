// tcgen05 / TMEM / TMA GEMM for the CLIP text tower (sm_100a only).
//
//   C[M,N] = A[M,K] . Bt[N,K]^T  (+ bias, activation, residual)      A, Bt bf16 K-major, fp32 accumulate
//
// This is the kernel behind every nn.Linear of the reference's text tower
// (/root/reference/src/open_clip/transformer.py:225 in_proj/out_proj, :233-235 c_fc/c_proj,
// /root/reference/src/open_clip/model.py:282 text_projection). nn.Linear weights are [out,in] row-major,
// i.e. already the K-major "B" operand, so no transpose is ever made.
//
// Structure (one persistent CTA per SM, 384 threads):
//   warp 0      TMA producer   : cp.async.bulk.tensor 128B-swizzled A (128x64) and B (256x64) tiles into a
//                                4-stage shared-memory ring, completion on mbarriers
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M128 N256 K16),
//                                accumulators in TMEM, double buffered (2 x 256 columns)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue       : tcgen05.ld 32 lanes x 32 columns -> registers -> fused bias / GELU / residual
//                                -> global store; overlaps the next tile's MMAs through the TMEM double buffer
// M may be a device-side value (packed token rows are only known on the device).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace leaf {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_THREADS = 384;
constexpr uint32_t GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;
constexpr uint32_t GEMM_B_BYTES = GEMM_BN * GEMM_BK * 2;
constexpr uint32_t GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr uint32_t GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

enum { EPI_BF16 = 0, EPI_BF16_ACT = 1, EPI_F32_RESIDUAL = 2, EPI_F32 = 3 };
enum { ACT_GELU_ERF = 0, ACT_QUICK_GELU = 1 };

struct GemmParams {
  int M;                  // rows (upper bound when m_dev != nullptr)
  const int* m_dev;       // optional device-side row count
  int N, K;
  const float* bias;      // [N] or nullptr
  void* C;                // bf16 or fp32 [M, ldc]
  int ldc;
  int act;
  uint32_t tx_bytes;      // bytes one pipeline stage receives (TMA boxes are clamped to small tensors)
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must end in a trap (reported as a CUDA error), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (it == 1024) t0 = clock64();
    if (it > 1024 && (it & 1023) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzled operand tile whose rows are 128 bytes (64 bf16): 8-row groups are 1024 B apart
// (SBO), LBO is unused by swizzled K-major layouts (encoded 1), descriptor version 1 (Blackwell), layout 2.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// nn.GELU (erf form): 0.5 x (1 + erf(x / sqrt 2)). erfc(z) = poly(t) exp(-z^2), t = 1/(1 + p z) (Abramowitz-Stegun
// 7.1.26, |error| <= 1.5e-7 on erf), evaluated on |x| so that the negative tail has no cancellation. ~14 FP32 ops
// + 2 MUFU per element: libdevice's erff costs ~39 instructions and made the fc1 epilogue the bottleneck (profiles/).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = fast_rcp(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  poly *= t;
  const float e = fast_ex2(x * x * -0.72134752044448170368f);            // exp(-x^2 / 2)
  const float h = 0.5f * x * poly * e;                                   // 0.5 x erfc(|x| / sqrt 2)
  return x >= 0.f ? x - h : h;
}
__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == ACT_QUICK_GELU) return x * fast_rcp(1.0f + fast_ex2(-1.702f * 1.4426950408889634f * x));   // x sigmoid(1.702 x)
  return gelu_erf(x);
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// fused epilogue of one 32-column chunk of one accumulator row (v = raw fp32 bits from TMEM)
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const uint32_t* v, int row, int col0) {
            float f[32];
  #pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (p.bias) {
  #pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (col0 + j < p.N) {
                  const float4 b = *reinterpret_cast<const float4*>(p.bias + col0 + j);
                  f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
                }
              }
            }
            if (EPI == EPI_BF16_ACT) {
  #pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = act_apply(f[j], p.act);
            }
            if (EPI == EPI_BF16 || EPI == EPI_BF16_ACT) {
              __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.C) + static_cast<size_t>(row) * p.ldc + col0;
  #pragma unroll
              for (int j = 0; j < 32; j += 8) {
                if (col0 + j < p.N) {
                  uint4 pk;
                  __nv_bfloat162 t0 = __floats2bfloat162_rn(f[j], f[j + 1]);
                  __nv_bfloat162 t1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
                  __nv_bfloat162 t2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
                  __nv_bfloat162 t3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                  pk.x = *reinterpret_cast<uint32_t*>(&t0);
                  pk.y = *reinterpret_cast<uint32_t*>(&t1);
                  pk.z = *reinterpret_cast<uint32_t*>(&t2);
                  pk.w = *reinterpret_cast<uint32_t*>(&t3);
                  *reinterpret_cast<uint4*>(out + j) = pk;
                }
              }
            } else {
              float* out = reinterpret_cast<float*>(p.C) + static_cast<size_t>(row) * p.ldc + col0;
  #pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (col0 + j < p.N) {
                  float4 o = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                  if (EPI == EPI_F32_RESIDUAL) {
                    const float4 r = *reinterpret_cast<const float4*>(out + j);
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                  }
                  *reinterpret_cast<float4*>(out + j) = o;
                }
              }
            }
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B needs 1024 B alignment
  const uint32_t bar_base = smem_base + GEMM_STAGES * GEMM_STAGE_BYTES;
  // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (GEMM_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * GEMM_STAGES + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + GEMM_STAGES * GEMM_STAGE_BYTES +
                                                                         8u * (2 * GEMM_STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int M = p.m_dev ? min(*p.m_dev, p.M) : p.M;
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (p.N + GEMM_BN - 1) / GEMM_BN;
  const int k_blocks = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);       // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * GEMM_STAGE_BYTES;
          const uint32_t sb = sa + GEMM_A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), p.tx_bytes);
          tma_load_2d(sa, &tmap_a, full_bar(stage), kb * GEMM_BK, m_blk * GEMM_BM);
          tma_load_2d(sb, &tmap_b, full_bar(stage), kb * GEMM_BK, n_blk * GEMM_BN);
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (a single thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, GEMM_BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * GEMM_BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * GEMM_STAGE_BYTES;
          const uint32_t sb = sa + GEMM_A_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));                 // frees the smem stage when these MMAs retire
          if (kb == k_blocks - 1) umma_commit(tfull_bar(acc));
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: 8 warps, quadrant = warp % 4 (TMEM lane window), column half = (warp - 4) / 4 =====
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    // residual epilogue: the fp32 tile it will read-modify-write is pulled into L2 one tile ahead, while the MMAs of
    // that tile are still running (the out-proj GEMM sits at the HBM/tensor ridge, profiles/r1)
    auto prefetch_residual = [&](int tile) {
      if (EPI != EPI_F32_RESIDUAL || tile >= total_tiles) return;
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int row = m_blk * GEMM_BM + quad * 32 + lane;
      if (row >= M) return;
      const int col0 = n_blk * GEMM_BN + half * (GEMM_BN / 2);
      const float* r = reinterpret_cast<const float*>(p.C) + static_cast<size_t>(row) * p.ldc + col0;
#pragma unroll
      for (int c = 0; c < GEMM_BN / 2; c += 32)
        if (col0 + c < p.N) prefetch_l2(r + c);
    };
    prefetch_residual(blockIdx.x);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      prefetch_residual(tile + gridDim.x);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * GEMM_BM + quad * 32 + lane;
      const bool row_ok = row < M;
#pragma unroll 1
      for (int c = 0; c < GEMM_BN / 2; c += 32) {
        const int col0 = n_blk * GEMM_BN + half * (GEMM_BN / 2) + c;
        if (col0 >= p.N) break;                          // warp-uniform
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>(acc * GEMM_BN + half * (GEMM_BN / 2) + c);
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        if (row_ok) epilogue_chunk<EPI>(p, v, row, col0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace leaf
