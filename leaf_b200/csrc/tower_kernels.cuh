// Non-GEMM kernels of the text tower and the TextFARE score (sm_100a).
// Token rows are PACKED: sequence i owns rows [cu[i], cu[i] + len[i]) where len[i] = argmax(ids)+1; positions
// after the pooled EOS row cannot influence it under the causal mask (transformer.py:758-764) and are skipped.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace leaf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of the row lengths -> cu[N+1]; also clamps len to [1, 77]. Single CTA (N <= ~10^5).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) scan_lengths_kernel(const int* __restrict__ len, int N, int* __restrict__ cu,
                                                            int* __restrict__ total_rows) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < N; base += 1024) {
    const int i = base + threadIdx.x;
    int v = (i < N) ? min(max(len[i], 1), 77) : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      warp_tot[lane] = winc - w;   // exclusive
    }
    __syncthreads();
    const int excl = carry + warp_tot[warp] + inc - v;
    if (i < N) cu[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    cu[N] = carry;
    *total_rows = carry;
  }
}

// ---------------------------------------------------------------------------------------------
// x[row,:] = token_embedding[id] + positional_embedding[pos]      (model.py:272-274), fp32
// one warp per packed row; grid = sequences
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_kernel(const int* __restrict__ tok, const int* __restrict__ cu, int N, int W,
                                                    const float* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                                    float* __restrict__ x) {
  const int seq = blockIdx.x;
  if (seq >= N) return;
  const int start = cu[seq], t = cu[seq + 1] - start;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int pos = warp; pos < t; pos += nwarp) {
    const int id = tok[seq * 77 + pos];
    const float4* e = reinterpret_cast<const float4*>(tok_emb + static_cast<size_t>(id) * W);
    const float4* pe = reinterpret_cast<const float4*>(pos_emb + static_cast<size_t>(pos) * W);
    float4* o = reinterpret_cast<float4*>(x + static_cast<size_t>(start + pos) * W);
    for (int c = lane; c < W / 4; c += 32) {
      float4 a = __ldg(e + c), b = __ldg(pe + c);
      o[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over fp32 rows -> bf16 (F.layer_norm, transformer.py:24-30), two-pass in registers.
// W = 128 * V4 floats... one warp per row, lane owns W/32 contiguous-by-4 elements. Rows from a device count.
// gather != nullptr: output row r reads input row gather[r] (EOS pooling, transformer.py:661).
// ---------------------------------------------------------------------------------------------
template <int VPL /* float4 per lane */>
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const float* __restrict__ x, const int* __restrict__ rows_dev,
                                                            int rows_max, const int* __restrict__ gather, int W,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, __nv_bfloat16* __restrict__ y) {
  const int rows = rows_dev ? min(*rows_dev, rows_max) : rows_max;
  const int lane = threadIdx.x & 31;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps_total) {
    const int src = gather ? gather[r] : r;
    const float4* in = reinterpret_cast<const float4*>(x + static_cast<size_t>(src) * W);
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      v[i] = in[lane + 32 * i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / static_cast<float>(W);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(W) + eps);
    uint2* out = reinterpret_cast<uint2*>(y + static_cast<size_t>(r) * W);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
      __nv_bfloat162 lo = __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y);
      __nv_bfloat162 hi = __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      out[lane + 32 * i] = pk;
    }
  }
}

// pooled row index of every sequence: cu[i] + len[i] - 1
__global__ void eos_rows_kernel(const int* __restrict__ cu, int N, int* __restrict__ eos_row) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) eos_row[i] = cu[i + 1] - 1;
}

// ---------------------------------------------------------------------------------------------
// causal attention over packed rows, head_dim 64 (nn.MultiheadAttention with the additive -inf mask,
// transformer.py:225,250-252,758-764; scale 1/sqrt(64)). One CTA per (sequence, head); K and V of the
// sequence staged in shared memory as fp32; one thread per query row with an online softmax in fp32.
// qkv: bf16 [rows, 3W] (q | k | v), out: bf16 [rows, W].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(96) attention_kernel(const __nv_bfloat16* __restrict__ qkv, const int* __restrict__ cu,
                                                       int W, __nv_bfloat16* __restrict__ out) {
  __shared__ float Ks[77][64];
  __shared__ float Vs[77][64];
  const int seq = blockIdx.x, head = blockIdx.y;
  const int start = cu[seq], t = cu[seq + 1] - start;
  const size_t ld = static_cast<size_t>(3) * W;
  // stage K, V: 8 bf16 (16 B) per thread-iteration
  for (int idx = threadIdx.x; idx < t * 8; idx += blockDim.x) {
    const int r = idx >> 3, c8 = (idx & 7) * 8;
    const __nv_bfloat16* base = qkv + (static_cast<size_t>(start + r)) * ld + head * 64 + c8;
    const uint4 kk = *reinterpret_cast<const uint4*>(base + W);
    const uint4 vv = *reinterpret_cast<const uint4*>(base + 2 * W);
    const __nv_bfloat162* kp = reinterpret_cast<const __nv_bfloat162*>(&kk);
    const __nv_bfloat162* vp = reinterpret_cast<const __nv_bfloat162*>(&vv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 kf = __bfloat1622float2(kp[j]);
      const float2 vf = __bfloat1622float2(vp[j]);
      Ks[r][c8 + 2 * j] = kf.x; Ks[r][c8 + 2 * j + 1] = kf.y;
      Vs[r][c8 + 2 * j] = vf.x; Vs[r][c8 + 2 * j + 1] = vf.y;
    }
  }
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= t) return;
  float q[64], o[64];
  {
    const __nv_bfloat16* qp = qkv + (static_cast<size_t>(start + i)) * ld + head * 64;
#pragma unroll
    for (int c = 0; c < 64; c += 8) {
      const uint4 qq = *reinterpret_cast<const uint4*>(qp + c);
      const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&qq);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(q2[j]);
        q[c + 2 * j] = f.x * 0.125f; q[c + 2 * j + 1] = f.y * 0.125f;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 64; ++c) o[c] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j <= i; ++j) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 64; ++c) s = fmaf(q[c], Ks[j][c], s);
    const float mn = fmaxf(m, s);
    const float corr = __expf(m - mn);
    const float pj = __expf(s - mn);
    l = l * corr + pj;
#pragma unroll
    for (int c = 0; c < 64; ++c) o[c] = fmaf(o[c], corr, pj * Vs[j][c]);
    m = mn;
  }
  const float inv = 1.f / l;
  __nv_bfloat16* op = out + (static_cast<size_t>(start + i)) * W + head * 64;
#pragma unroll
  for (int c = 0; c < 64; c += 8) {
    uint4 pk;
    __nv_bfloat162 t0 = __floats2bfloat162_rn(o[c] * inv, o[c + 1] * inv);
    __nv_bfloat162 t1 = __floats2bfloat162_rn(o[c + 2] * inv, o[c + 3] * inv);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(o[c + 4] * inv, o[c + 5] * inv);
    __nv_bfloat162 t3 = __floats2bfloat162_rn(o[c + 6] * inv, o[c + 7] * inv);
    pk.x = *reinterpret_cast<uint32_t*>(&t0);
    pk.y = *reinterpret_cast<uint32_t*>(&t1);
    pk.z = *reinterpret_cast<uint32_t*>(&t2);
    pk.w = *reinterpret_cast<uint32_t*>(&t3);
    *reinterpret_cast<uint4*>(op + c) = pk;
  }
}

// F.normalize(x, dim=-1) in place (model.py:284): x / max(||x||, 1e-12). One warp per row.
__global__ void __launch_bounds__(256) l2_normalize_kernel(float* __restrict__ f, int N, int E) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= N) return;
  float* row = f + static_cast<size_t>(r) * E;
  float s = 0.f;
  for (int c = lane; c < E; c += 32) s += row[c] * row[c];
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  for (int c = lane; c < E; c += 32) row[c] *= inv;
}

// ---------------------------------------------------------------------------------------------
// K3: score[b,j] and per-sample argmax (utils_attacks.py:332-348, :370-386, :393).
// One CTA per sample; each warp reduces whole candidates (float4 loads), then a CTA-level argmax with the
// FIRST maximal index winning, as torch.argmax does. objective: 0 l2, 1 negl2, 2 sim, 3 dissim.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) score_argmax_kernel(const float* __restrict__ feat, const float* __restrict__ anchor,
                                                          int n, int E, int objective, float* __restrict__ loss_out,
                                                          int* __restrict__ best_out, float* __restrict__ best_feat_out) {
  extern __shared__ float sc[];     // [n] scores
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  __shared__ int best_s;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const float4* a4 = reinterpret_cast<const float4*>(anchor + static_cast<size_t>(b) * E);
  for (int j = warp; j < n; j += nwarp) {
    const float4* f4 = reinterpret_cast<const float4*>(feat + (static_cast<size_t>(b) * n + j) * E);
    float s = 0.f;
    for (int c = lane; c < E / 4; c += 32) {
      const float4 f = f4[c], a = __ldg(a4 + c);
      if (objective <= 1) {
        const float dx = f.x - a.x, dy = f.y - a.y, dz = f.z - a.z, dw = f.w - a.w;
        s += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      } else {
        s += (f.x * a.x + f.y * a.y) + (f.z * a.z + f.w * a.w);
      }
    }
    s = warp_sum(s);
    if (objective == 1 || objective == 3) s = -s;
    if (lane == 0) {
      sc[j] = s;
      if (loss_out) loss_out[static_cast<size_t>(b) * n + j] = s;
    }
  }
  __syncthreads();
  // argmax, first index on ties; NaN never wins over a number (torch treats NaN as max - not reachable here)
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float v = sc[j];
    if (v > bv || (v == bv && j < bi)) { bv = v; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nwarp; ++w)
      if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) { bv = red_v[w]; bi = red_i[w]; }
    if (bi == 0x7fffffff) bi = 0;
    best_s = bi;
    best_out[b] = bi;
  }
  __syncthreads();
  if (best_feat_out) {
    const float* src = feat + (static_cast<size_t>(b) * n + best_s) * E;
    float* dst = best_feat_out + static_cast<size_t>(b) * E;
    for (int c = threadIdx.x; c < E; c += blockDim.x) dst[c] = src[c];
  }
}

// fp32 -> bf16 cast (weight refresh), optional transpose for text_projection [W,E] -> [E,W]
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst + i) = pk;
  }
}
__global__ void cast_bf16_transpose_kernel(const float* __restrict__ src /*[R,C]*/, __nv_bfloat16* __restrict__ dst /*[C,R]*/,
                                           int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? src[static_cast<size_t>(r) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) dst[static_cast<size_t>(c) * R + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}
__global__ void copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) dst[i] = src[i];
}

}  // namespace leaf
