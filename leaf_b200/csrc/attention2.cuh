// Causal attention over packed rows, second form: every warp streams its (sequence, head) items through a private
// shared-memory ring that is filled ASYNCHRONOUSLY (cp.async, 16 B per lane) several tiles ahead of the tensor-core work.
//
// Why: the register-fed kernel (tower_kernels.cuh::attention_kernel) issues a (sequence, head)'s loads, waits a DRAM
// round trip, computes, stores, and only then starts the next item - at 16 warps per SM (128 registers) that leaves the
// memory system idle most of the time: 0.46 of the HBM copy rate, "pure latency" (ncu r45: occupancy 23 %, issue 39 %,
// DRAM bytes = algorithmic bytes). Here the loads of the NEXT elements - across key tiles, query tiles and items - are in
// flight while the current one is multiplied: a warp keeps AT2_LOOK elements (up to 4 KB each) outstanding at all times.
//
// Stream of one warp: for every item (sequence, head) it owns, for every 16-row query tile: [Q, KV0, KV1, ...]; a Q element
// is the head's slice of 16 query rows (2 KB), a KV element the K and the V slices of 16 key rows (2 x 2 KB); one element =
// one ring slot = one cp.async group. The producer cursor runs AT2_LOOK elements ahead of the consumer; both walk the same
// enumeration. Rows are addressed one by one (meta's shared-prefix indirection: keys [0, p) come from the base sequence's
// rows), which cp.async can do and a TMA box cannot. Tiles are XOR-swizzled (16-byte chunk ^= row & 7) so that cp.async
// writes and ldmatrix reads are bank-conflict free. Fragments come from ldmatrix (.trans for V); softmax is the online
// form (running row maximum, O rescaled per key tile), so registers do not grow with the sequence length. The output
// tile is staged through the V half of the last slot and written as whole 128-byte head rows.
//
// The first version of this kernel (one element per K / V tile, general cursor lambdas) hid the latency (ncu: long-scoreboard
// 0.29 per issue) but executed 1 936 instructions per item, 3 % of them HMMA, and lost to the register-fed kernel (312 vs
// 233 us); this one is written for instruction count: one address computation serves a key row's K and V copies, ldmatrix /
// staging addresses are one XOR away from two per-lane constants, the causal mask and the rescale are skipped for tiles
// that do not need them.
//
// Every output row is a function of its own query row and the key/value rows it sees, visited in the same order whatever the
// packing: results are bit-identical with and without shared prefixes / duplicate elimination / last-row pruning.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tower_kernels.cuh"

namespace leaf {

constexpr int AT2_WARPS = 8;
constexpr int AT2_LOOK = 3;                                          // elements in flight per warp
constexpr int AT2_SLOT = 4096;                                       // K tile | V tile (a Q element uses the first half)
constexpr int AT2_WARP_SMEM = AT2_LOOK * AT2_SLOT;
constexpr int AT2_SMEM_BYTES = AT2_WARPS * AT2_WARP_SMEM + 128;      // + alignment slack

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(AT2_WARPS * 32, 2) attention2_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                        const int4* __restrict__ meta, int n_seq, int heads,
                                                                        int W, __nv_bfloat16* __restrict__ out, int last_only) {
  extern __shared__ uint8_t at2_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const uint32_t ring = ((static_cast<uint32_t>(__cvta_generic_to_shared(at2_smem)) + 127u) & ~127u) + warp * AT2_WARP_SMEM;
  const int n_items = n_seq * heads;
  const int stride = gridDim.x * AT2_WARPS;
  const uint32_t ld = 3u * W;                               // element offsets fit 32 bits (launch_attention checks rows * 3 W < 2^32)
  const float sl2 = 0.125f * 1.4426950408889634f;          // 1/sqrt(64) * log2(e)
  // cp.async / staging: this lane moves 16-byte chunk `ch` of rows 4 i + lr; (row & 7) = lr ^ 4 (i & 1)
  const int lr = lane >> 3, ch = lane & 7;
  const uint32_t dA = lr * 128 + ((ch ^ lr) << 4), dB = lr * 128 + ((ch ^ lr ^ 4) << 4);
  // ldmatrix: this lane addresses row ri of matrix mi; (chunk ^ row & 7) << 4 is one XOR with k-step * 32 away from these
  const int mi = lane >> 3, ri = lane & 7;
  const uint32_t offQV = ((mi & 1) * 8 + ri) * 128 + ((((mi >> 1) ^ ri) & 1) << 4) + ((ri & 6) << 4);   // Q (A operand) and V (.trans)
  const uint32_t offK = ((mi >> 1) * 8 + ri) * 128 + ((((mi & 1) ^ ri) & 1) << 4) + ((ri & 6) << 4);    // K (B operand of Q.K^T)
  const uint32_t offO = g * 128 + (g << 4) + 4 * c;         // output staging: row g, chunk i -> ^ (i << 4); row g + 8: + 1024

  // ---- producer cursor (AT2_LOOK elements ahead) ----
  int p_item = blockIdx.x * AT2_WARPS + warp;
  int p_own = 0, p_t1 = 0, p_p = 0, p_base = 0, p_nq1 = 0, p_q0 = 0, p_nkt = 0, p_idx = 0;
  const __nv_bfloat16 *p_q = qkv, *p_k = qkv, *p_v = qkv;   // this lane's 16-byte chunk of the head's Q / K / V slice of row 0
  bool pv;
  auto p_load_item = [&]() {                                // warp-uniform; skips duplicates (no own rows)
    while (p_item < n_items) {
      const int seq = p_item / heads;
      const int4 mt = __ldg(meta + seq);
      if (mt.y > mt.z) {
        p_own = mt.x; p_t1 = mt.y - 1; p_p = mt.z; p_base = mt.w; p_nq1 = mt.y - mt.z - 1;
        p_q = qkv + (p_item - seq * heads) * 64 + ch * 8;
        p_k = p_q + W;
        p_v = p_k + W;
        p_q0 = last_only ? (p_nq1 & ~15) : 0;
        p_idx = 0;
        p_nkt = (min(p_t1, p_p + p_q0 + 15) >> 4) + 1;
        return true;
      }
      p_item += stride;
    }
    return false;
  };
  auto p_advance = [&]() {
    if (++p_idx <= p_nkt) return true;
    p_idx = 0;
    p_q0 += 16;
    if (p_q0 > p_nq1) { p_item += stride; return p_load_item(); }
    p_nkt = (min(p_t1, p_p + p_q0 + 15) >> 4) + 1;
    return true;
  };
  auto p_issue = [&](uint32_t dst) {
    if (p_idx == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t row = p_own + min(p_q0 + 4 * i + lr, p_nq1);
        cp_async16(dst + i * 512 + ((i & 1) ? dB : dA), p_q + row * ld);
      }
    } else {
      const int j0 = (p_idx - 1) * 16 + lr;
      const int delta = p_own - p_p;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = min(j0 + 4 * i, p_t1);
        const uint32_t off = static_cast<uint32_t>((j < p_p ? p_base : delta) + j) * ld;
        cp_async16(dst + i * 512 + ((i & 1) ? dB : dA), p_k + off);
        cp_async16(dst + 2048 + i * 512 + ((i & 1) ? dB : dA), p_v + off);
      }
    }
  };

  // ---- consumer cursor ----
  int c_item = p_item, c_seq = 0, c_head = 0, c_own = 0, c_t1 = 0, c_p = 0, c_nq = 0, c_q0 = 0;
  auto c_load_item = [&]() {
    while (c_item < n_items) {
      c_seq = c_item / heads;
      const int4 mt = __ldg(meta + c_seq);
      if (mt.y > mt.z) {
        c_head = c_item - c_seq * heads;
        c_own = mt.x; c_t1 = mt.y - 1; c_p = mt.z; c_nq = mt.y - mt.z;
        c_q0 = last_only ? ((c_nq - 1) & ~15) : 0;
        return true;
      }
      c_item += stride;
    }
    return false;
  };

  pv = p_load_item();
  bool cv = c_load_item();
#pragma unroll
  for (int i = 0; i < AT2_LOOK; ++i) {
    if (pv) { p_issue(ring + i * AT2_SLOT); pv = p_advance(); }
    cp_async_commit();
  }
  uint32_t slot = ring;                                     // address of the slot the consumer reads next
  auto release = [&]() {
    __syncwarp();                                           // every lane has read the slot (ldmatrix results are in registers)
    if (pv) { p_issue(slot); pv = p_advance(); }
    cp_async_commit();
    slot += AT2_SLOT;
    if (slot == ring + AT2_LOOK * AT2_SLOT) slot = ring;
  };

  while (cv) {
    // ---- Q tile -> A fragments ----
    uint32_t qf[4][4];
    cp_async_wait<AT2_LOOK - 1>();
    __syncwarp();
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4((slot + offQV) ^ (ks << 5), qf[ks]);
    release();
    const int pos0 = c_p + min(c_q0 + g, c_nq - 1), pos1 = c_p + min(c_q0 + g + 8, c_nq - 1);
    const int nkt = (min(c_t1, c_p + c_q0 + 15) >> 4) + 1;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    for (int kt = 0; kt < nkt; ++kt) {
      cp_async_wait<AT2_LOOK - 1>();
      __syncwarp();
      // ---- S = Q.K^T for 16 keys ----
      float sc[2][4];
      sc[0][0] = sc[0][1] = sc[0][2] = sc[0][3] = sc[1][0] = sc[1][1] = sc[1][2] = sc[1][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t b[4];
        ldsm_x4((slot + offK) ^ (ks << 5), b);
        mma_bf16_16816(sc[0], qf[ks], b[0], b[1]);
        mma_bf16_16816(sc[1], qf[ks], b[2], b[3]);
      }
      // ---- causal mask (only tiles that reach past the tile's first query position) ----
      if (kt * 16 + 15 > c_p + c_q0) {
        const int lim0 = pos0 - kt * 16 - 2 * c, lim1 = pos1 - kt * 16 - 2 * c;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (nt * 8 + e > lim0) sc[nt][e] = -INFINITY;
            if (nt * 8 + e > lim1) sc[nt][2 + e] = -INFINITY;
          }
      }
      // ---- running maximum, P = exp2((S - max) * scale) ----
      float t0 = fmaxf(fmaxf(sc[0][0], sc[0][1]), fmaxf(sc[1][0], sc[1][1]));
      float t1 = fmaxf(fmaxf(sc[0][2], sc[0][3]), fmaxf(sc[1][2], sc[1][3]));
      t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 1));
      t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 1));
      t0 = fmaxf(t0, __shfl_xor_sync(0xffffffffu, t0, 2));
      t1 = fmaxf(t1, __shfl_xor_sync(0xffffffffu, t1, 2));
      const float n0 = fmaxf(m0, t0), n1 = fmaxf(m1, t1);   // finite from the first tile on: key 0 is visible to every query
      const float nm0 = -n0 * sl2, nm1 = -n1 * sl2;
      uint32_t pf[4];
      float s0, s1;
      {
        const float p00 = att_ex2(fmaf(sc[0][0], sl2, nm0)), p01 = att_ex2(fmaf(sc[0][1], sl2, nm0));
        const float p10 = att_ex2(fmaf(sc[0][2], sl2, nm1)), p11 = att_ex2(fmaf(sc[0][3], sl2, nm1));
        const float q00 = att_ex2(fmaf(sc[1][0], sl2, nm0)), q01 = att_ex2(fmaf(sc[1][1], sl2, nm0));
        const float q10 = att_ex2(fmaf(sc[1][2], sl2, nm1)), q11 = att_ex2(fmaf(sc[1][3], sl2, nm1));
        s0 = (p00 + p01) + (q00 + q01);
        s1 = (p10 + p11) + (q10 + q11);
        pf[0] = pack_bf16x2(p00, p01); pf[1] = pack_bf16x2(p10, p11);
        pf[2] = pack_bf16x2(q00, q01); pf[3] = pack_bf16x2(q10, q11);
      }
      if (kt == 0) {                                        // nothing accumulated yet: no rescale
        l0 = s0; l1 = s1;
      } else {
        // (m - n) * scale, NOT fma(m, scale, -n * scale): a key tile that is fully masked for a row must rescale it by exactly 1
        // (which tiles a row meets beyond its own keys depends on the packing; results must not)
        const float r0 = att_ex2((m0 - n0) * sl2), r1 = att_ex2((m1 - n1) * sl2);
        l0 = fmaf(l0, r0, s0);
        l1 = fmaf(l1, r1, s1);
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] *= r0; o[i][1] *= r0; o[i][2] *= r1; o[i][3] *= r1; }
      }
      m0 = n0; m1 = n1;
      // ---- O += P.V ----
#pragma unroll
      for (int ip = 0; ip < 4; ++ip) {
        uint32_t b[4];
        ldsm_x4_trans((slot + 2048 + offQV) ^ (ip << 5), b);
        mma_bf16_16816(o[2 * ip], pf, b[0], b[1]);
        mma_bf16_16816(o[2 * ip + 1], pf, b[2], b[3]);
      }
      if (kt + 1 < nkt) release();
    }
    // ---- normalise, stage the 16 x 64 tile in the V half of the last slot, write whole 128-byte head rows ----
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    __syncwarp();                                           // every lane's V fragments are in registers
    const uint32_t stage = slot + 2048;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sts32((stage + offO) ^ (i << 4), pack_bf16x2(o[i][0] * inv0, o[i][1] * inv0));
      sts32((stage + 1024 + offO) ^ (i << 4), pack_bf16x2(o[i][2] * inv1, o[i][3] * inv1));
    }
    __syncwarp();
    {
      __nv_bfloat16* ob = out + c_head * 64 + ch * 8;
      const int left = c_nq - c_q0;                          // rows of this tile that exist
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 4 + lr;
        const uint4 v = lds128(stage + i * 512 + ((i & 1) ? dB : dA));
        if (last_only) {
          if (r == left - 1) *reinterpret_cast<uint4*>(ob + static_cast<long long>(c_seq) * W) = v;
        } else if (r < left) {
          *reinterpret_cast<uint4*>(ob + static_cast<long long>(c_own + c_q0 + r) * W) = v;
        }
      }
    }
    release();                                              // the last KV slot, staging included
    c_q0 += 16;
    if (c_q0 >= c_nq) { c_item += stride; cv = c_load_item(); }
  }
  cp_async_wait<0>();
}

}  // namespace leaf
