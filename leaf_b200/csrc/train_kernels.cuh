// K4: kernels of the backward pass of the selected adversarial batch (the FARE-style update,
// /root/reference/utils_AT.py:312-337: encode_text(adv) in train mode, mse(...).sum(-1).mean(), backward).
// All matrix products (dgrad and wgrad) run on the tcgen05 GEMM (gemm2_sm100.cuh) as K-major TN products; the kernels
// here are the element-wise / per-row pieces around them and the operand transposes the wgrad products need.
// The batch is the B winners only (~3 % of the step's FLOPs), so these kernels are written for clarity and exact
// fp32 math, not for the roofline.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tower_kernels.cuh"

namespace leaf {

// ---- activation forward / backward (fc1 output u is saved pre-activation) -----------------------------------------
__device__ __forceinline__ float act_fwd_exact(float u, int act) {
  if (act == 1) return u / (1.f + expf(-1.702f * u));
  return 0.5f * u * (1.f + erff(u * 0.70710678118654752440f));
}
__device__ __forceinline__ float act_grad_exact(float u, int act) {
  if (act == 1) {
    const float s = 1.f / (1.f + expf(-1.702f * u));
    return s + 1.702f * u * s * (1.f - s);
  }
  const float cdf = 0.5f * (1.f + erff(u * 0.70710678118654752440f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * u * u);
  return cdf + u * pdf;
}
__global__ void act_fwd_kernel(const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ g, size_t n, int act) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    g[i] = __float2bfloat16_rn(act_fwd_exact(__bfloat162float(u[i]), act));
}
// du = dg * act'(u)   (dg fp32 from the dgrad GEMM, du bf16 operand of the next GEMMs)
__global__ void act_bwd_kernel(const float* __restrict__ dg, const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ du,
                               size_t n, int act) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    du[i] = __float2bfloat16_rn(dg[i] * act_grad_exact(__bfloat162float(u[i]), act));
}
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// ---- dst[C, Rp] (bf16) = src[R, C]^T, rows r >= R zero-filled (Rp = R rounded up to 8: TMA row pitch) --------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void transpose_bf16_kernel(const T* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int C, int Rp) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? to_f32<T>(src[static_cast<size_t>(r) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < Rp) dst[static_cast<size_t>(c) * Rp + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

// ---- dst[C] += column sums of src[R, C] (bias gradients) -------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ src, int R, int C, int ld, float* __restrict__ dst) {
  __shared__ float part[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C)
    for (int r = blockIdx.y * 8 + w; r < R; r += gridDim.y * 8) s += to_f32<T>(src[static_cast<size_t>(r) * ld + c]);
  part[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    atomicAdd(dst + c, t);
  }
}

// ---- LayerNorm backward: dx[row] (+)= rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)); dgamma += dy*xhat; dbeta += dy
// One warp per row. gather != nullptr: row r reads x[gather[r]] and WRITES dx[gather[r]] (final LN on the pooled rows).
// accumulate: dx += (residual branch) instead of dx = .
template <int VPL>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const int* __restrict__ gather, int rows, int W,
                                                           const float* __restrict__ gamma, float eps, float* __restrict__ dx,
                                                           int accumulate, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  float4 dg_acc[VPL], db_acc[VPL], g[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    dg_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
  }
  const float invW = 1.f / static_cast<float>(W);
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps_total) {
    const int xr = gather ? gather[r] : r;
    const float4* xin = reinterpret_cast<const float4*>(x + static_cast<size_t>(xr) * W);
    const float4* din = reinterpret_cast<const float4*>(dy + static_cast<size_t>(r) * W);
    float4 xv[VPL], dv[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xv[i] = xin[lane + 32 * i];
      dv[i] = din[lane + 32 * i];
      s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
    const float mean = warp_sum(s) * invW;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * invW + eps);
    float s1 = 0.f, s2 = 0.f;                 // sum(g*dy), sum(g*dy*xhat)
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;      // xhat
      dg_acc[i].x += dv[i].x * xv[i].x; dg_acc[i].y += dv[i].y * xv[i].y;
      dg_acc[i].z += dv[i].z * xv[i].z; dg_acc[i].w += dv[i].w * xv[i].w;
      db_acc[i].x += dv[i].x; db_acc[i].y += dv[i].y; db_acc[i].z += dv[i].z; db_acc[i].w += dv[i].w;
      dv[i].x *= g[i].x; dv[i].y *= g[i].y; dv[i].z *= g[i].z; dv[i].w *= g[i].w;   // g*dy
      s1 += (dv[i].x + dv[i].y) + (dv[i].z + dv[i].w);
      s2 += (dv[i].x * xv[i].x + dv[i].y * xv[i].y) + (dv[i].z * xv[i].z + dv[i].w * xv[i].w);
    }
    s1 = warp_sum(s1) * invW;
    s2 = warp_sum(s2) * invW;
    float4* out = reinterpret_cast<float4*>(dx + static_cast<size_t>(xr) * W);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 o;
      o.x = rstd * (dv[i].x - s1 - xv[i].x * s2);
      o.y = rstd * (dv[i].y - s1 - xv[i].y * s2);
      o.z = rstd * (dv[i].z - s1 - xv[i].z * s2);
      o.w = rstd * (dv[i].w - s1 - xv[i].w * s2);
      if (accumulate) {
        const float4 p = out[lane + 32 * i];
        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      }
      out[lane + 32 * i] = o;
    }
  }
  // CTA-level reduction (warps take turns on a shared accumulator), then ONE atomic per column and CTA
  __shared__ float red[2][VPL * 128];
  for (int c = threadIdx.x; c < 2 * VPL * 128; c += blockDim.x) (&red[0][0])[c] = 0.f;
  __syncthreads();
  for (int w = 0; w < (blockDim.x >> 5); ++w) {
    if ((threadIdx.x >> 5) == w) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        float* a = &red[0][(lane + 32 * i) * 4];
        float* b = &red[1][(lane + 32 * i) * 4];
        a[0] += dg_acc[i].x; a[1] += dg_acc[i].y; a[2] += dg_acc[i].z; a[3] += dg_acc[i].w;
        b[0] += db_acc[i].x; b[1] += db_acc[i].y; b[2] += db_acc[i].z; b[3] += db_acc[i].w;
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    atomicAdd(dgamma + c, red[0][c]);
    atomicAdd(dbeta + c, red[1][c]);
  }
}

// ---- causal attention backward, one CTA per (sequence, head), fp32 in shared memory -------------------------------------
// qkv bf16 [rows,3W], o bf16 [rows,W] (forward output), dout fp32 [rows,W] -> dqkv bf16 [rows,3W].
//   P = softmax(scale Q K^T + causal), D_i = sum_d dO_id O_id, dP = dO V^T, dS = P o (dP - D),
//   dV = P^T dO, dQ = scale dS K, dK = scale dS^T Q.
constexpr int ATTB_THREADS = 256;
__host__ __device__ constexpr int attb_smem_bytes(int T) { return (4 * T * 65 + 2 * T * (T + 1) + T) * 4; }

// T = longest sequence of the batch (sizes the shared-memory arrays, so short captions fit several CTAs per SM)
__global__ void __launch_bounds__(ATTB_THREADS) attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                     const __nv_bfloat16* __restrict__ o,
                                                                     const float* __restrict__ dout, const int4* __restrict__ meta,
                                                                     int W, int T, __nv_bfloat16* __restrict__ dqkv) {
  extern __shared__ float sm[];
  float* Q = sm;
  float* K = Q + T * 65;
  float* V = K + T * 65;
  float* dO = V + T * 65;
  float* P = dO + T * 65;
  float* dS = P + T * (T + 1);
  float* Dv = dS + T * (T + 1);
  const int TP = T + 1;
  const int seq = blockIdx.x, head = blockIdx.y;
  const int4 mt = meta[seq];
  const int row0 = mt.x, t = min(mt.y, T);             // training batches are packed without prefix sharing (p = 0)
  const size_t ld = static_cast<size_t>(3) * W;
  for (int idx = threadIdx.x; idx < t * 64; idx += blockDim.x) {
    const int r = idx >> 6, d = idx & 63;
    const __nv_bfloat16* b = qkv + static_cast<size_t>(row0 + r) * ld + head * 64 + d;
    Q[r * 65 + d] = __bfloat162float(b[0]);
    K[r * 65 + d] = __bfloat162float(b[W]);
    V[r * 65 + d] = __bfloat162float(b[2 * W]);
    dO[r * 65 + d] = dout[static_cast<size_t>(row0 + r) * W + head * 64 + d];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < t; i += blockDim.x) {  // D_i
    float s = 0.f;
    for (int d = 0; d < 64; ++d) s += dO[i * 65 + d] * __bfloat162float(o[static_cast<size_t>(row0 + i) * W + head * 64 + d]);
    Dv[i] = s;
  }
  for (int idx = threadIdx.x; idx < t * t; idx += blockDim.x) {   // scores and dP
    const int i = idx / t, j = idx - i * t;
    float s = 0.f, dp = 0.f;
    if (j <= i) {
      for (int d = 0; d < 64; ++d) { s += Q[i * 65 + d] * K[j * 65 + d]; dp += dO[i * 65 + d] * V[j * 65 + d]; }
      s *= 0.125f;
    } else {
      s = -INFINITY;
    }
    P[i * TP + j] = s;
    dS[i * TP + j] = dp;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < t; i += blockDim.x) {  // row softmax, then dS
    float m = -INFINITY;
    for (int j = 0; j <= i; ++j) m = fmaxf(m, P[i * TP + j]);
    float l = 0.f;
    for (int j = 0; j <= i; ++j) { const float e = expf(P[i * TP + j] - m); P[i * TP + j] = e; l += e; }
    const float inv = 1.f / l, di = Dv[i];
    for (int j = 0; j < t; ++j) {
      const float pj = j <= i ? P[i * TP + j] * inv : 0.f;
      P[i * TP + j] = pj;
      dS[i * TP + j] = pj * (dS[i * TP + j] - di);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < t * 64; idx += blockDim.x) {
    const int r = idx >> 6, d = idx & 63;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j <= r; ++j) dq += dS[r * TP + j] * K[j * 65 + d];
    for (int i = r; i < t; ++i) { dk += dS[i * TP + r] * Q[i * 65 + d]; dv += P[i * TP + r] * dO[i * 65 + d]; }
    __nv_bfloat16* b = dqkv + static_cast<size_t>(row0 + r) * ld + head * 64 + d;
    b[0] = __float2bfloat16_rn(dq * 0.125f);
    b[W] = __float2bfloat16_rn(dk * 0.125f);
    b[2 * W] = __float2bfloat16_rn(dv);
  }
}

// ---- embedding backward: dtok[id] += dx[row], dpos[pos] += dx[row] -------------------------------------------------------
__global__ void __launch_bounds__(256) embed_bwd_kernel(const int* __restrict__ tok, const int4* __restrict__ meta, int N, int W,
                                                        const float* __restrict__ dx, float* __restrict__ dtok,
                                                        float* __restrict__ dpos) {
  const int seq = blockIdx.x;
  if (seq >= N) return;
  const int4 mt = meta[seq];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int pos = mt.z + warp; pos < mt.y; pos += nwarp) {
    const int id = tok[seq * 77 + pos];
    const float* g = dx + static_cast<size_t>(mt.x + pos - mt.z) * W;
    for (int c = lane; c < W; c += 32) {
      const float v = g[c];
      atomicAdd(dtok + static_cast<size_t>(id) * W + c, v);
      atomicAdd(dpos + static_cast<size_t>(pos) * W + c, v);
    }
  }
}

// ---- AdamW over the tower's flat parameter buffer (torch.optim.AdamW semantics, decoupled weight decay;
// train_AT_text_only.py:326-341: gains / biases / LayerNorm parameters form a group with weight_decay = 0 and are laid
// out first, elements [0, n_nodecay)). g is multiplied by grad_scale first (1/accum_freq, or the clipping
// coefficient). bc1 = 1 - beta1^t, bc2s = sqrt(1 - beta2^t). HBM bound: 28 B per element.
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, size_t n, size_t n_nodecay, float lr, float beta1,
                                                    float beta2, float eps, float wd, float bc1, float bc2s, float grad_scale) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 4;
  for (size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    float4 pv = *reinterpret_cast<const float4*>(p + i), gv = *reinterpret_cast<const float4*>(g + i);
    float4 mv = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
    float* pp = &pv.x; float* gg = &gv.x; float* mm = &mv.x; float* vs = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = gg[j] * grad_scale;
      if (i + j >= n_nodecay) pp[j] *= 1.f - lr * wd;
      mm[j] = beta1 * mm[j] + (1.f - beta1) * gj;              // lerp(m, g, 1 - beta1)
      vs[j] = beta2 * vs[j] + (1.f - beta2) * gj * gj;
      const float denom = sqrtf(vs[j]) / bc2s + eps;
      pp[j] -= (lr / bc1) * (mm[j] / denom);
    }
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
  }
}

// sum of squares of a flat fp32 buffer (gradient-norm clipping, utils_AT.py:349-357): out[0] += sum(g^2)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out) {
  __shared__ float part[8];
  float s = 0.f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 4;
  for (size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(g + i);
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i];
    atomicAdd(out, t);
  }
}

// dst[r, :] += src[r, :] (fp32), used to add the attention/MLP branch gradient into the running dx
__global__ void add_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] += src[i];
}

// dst[W,E] += src[E,W]^T (HF projection layout -> open_clip layout or back), fp32
__global__ void add_transposed_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? src[static_cast<size_t>(r) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) dst[static_cast<size_t>(c) * R + r] += tile[threadIdx.x][i];
  }
}

}  // namespace leaf
