// Kernels of the on-device `--constrain` filter (constrain_core.cuh): one THREAD per sentence. The work per sentence is
// a few thousand dependent byte operations over ~100 bytes; 12 928 sentences per phase keep every SM busy and the whole
// pass costs less than the tokenizer kernel. Buffers live in local memory (L1-resident at this footprint).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "constrain_core.cuh"
#include "k1_core.cuh"

namespace leaf {

struct CnArgs {
  const uint8_t* caps;
  const int32_t* cap_off;
  int B, n;
  const int32_t* pos;
  const int32_t* chr;
  const int32_t* sel;
  int32_t* count_out;     // [B*n + B]: candidates, then the B current sentences
  int32_t* status_out;
};

__global__ void __launch_bounds__(64) constrain_count_kernel(const CnTables T, const CnArgs a) {
  const int n_cand = a.B * a.n;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_cand + a.B) return;
  const bool is_base = r >= n_cand;
  const int b = is_base ? r - n_cand : r / a.n;
  const int off = a.cap_off[b];
  int len = a.cap_off[b + 1] - off;
  int flags = 0;
  uint8_t text[CN_MAX_TEXT + 8];
  uint8_t buf_a[CN_BUF], buf_b[CN_BUF];
  if (len > CN_MAX_TEXT - 1) { flags |= CN_FLAG_TOO_LONG; len = 0; }
  const uint8_t* src = a.caps + off;
  int m;
  if (is_base) {
    for (int i = 0; i < len; ++i) text[i] = src[i];
    m = len;
  } else {
    const int z = a.sel ? a.pos[b * a.n + a.sel[b]] : a.pos[r];
    const int c = a.chr[r];
    if (z < 0 || z > 2 * len) {
      flags |= K1_FLAG_TOO_LONG;
      for (int i = 0; i < len; ++i) text[i] = src[i];
      m = len;
    } else {
      m = k1_apply_edit(src, len, z, c, text);
    }
  }
  for (int i = 0; i < m; ++i) {                       // str.lower() on the ASCII domain
    uint8_t c = text[i];
    if (c >= 0x80) flags |= K1_FLAG_NON_ASCII;
    if (c >= 'A' && c <= 'Z') c += 32;
    text[i] = c;
  }
  a.count_out[r] = cn_count_words(T, text, m, buf_a, buf_b, flags);
  if (flags && a.status_out) atomicOr(a.status_out, flags);
}

// valid[b,j] = count(candidate) < count(current sentence)     (utils_attacks.py:143)
__global__ void constrain_valid_kernel(const int32_t* __restrict__ count, int B, int n, uint8_t* __restrict__ valid) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * n) return;
  valid[r] = count[r] < count[B * n + r / n] ? 1 : 0;
}

}  // namespace leaf
