// K1 kernel: candidate expansion + CLIP BPE tokenization, one warp per candidate (sm_100a). Captions are UTF-8 with code
// points <= U+024F (k1_core.cuh: domain); edit positions count code points, as Python's str does.
//   lanes 0..31 copy the caption into shared memory with 16-byte loads when the source is aligned,
//   lane 0 applies the edit, unescapes, cleans and splits (serial, a few hundred byte operations),
//   lanes take regex pieces round-robin and run the BPE merge loop on them (merge ranks from an L2-resident
//   1 MiB hash table: the 48 894-entry table does not fit a CTA's shared memory next to the scratch),
//   lane 0 assembles the 77-slot row; all lanes store it (coalesced int32).
// The scalar pieces are k1_core.cuh, which the CPU test-suite pins against the reference's tokenizer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "k1_core.cuh"

namespace leaf {

constexpr int K1_WARPS_PER_CTA = 3;          // default variant: captions up to 1000 bytes, 3 warps x 13 KB of scratch per CTA
constexpr int K1_LONG_TEXT = 3584;           // long variant: captions up to 3560 bytes, one warp (46 KB of scratch) per CTA

struct K1Args {
  const uint8_t* caps;
  const int32_t* cap_off;
  int B, n;
  const int32_t* pos;
  const int32_t* chr;
  const int32_t* sel;
  const uint8_t* valid;
  int32_t* tok_out;
  int32_t* len_out;
  int32_t* base_out;      // [R] index of the row holding the sample's unedited caption, -1 for those rows
  int32_t* status_out;
  int hf_mode;            // 1: HF CLIPTokenizer semantics (leaf_set_tokenizer_mode)
};

template <int MAXT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k1_expand_tokenize_kernel(const K1Tables T, const K1Args a) {
  __shared__ __align__(16) uint8_t s_src[WARPS][MAXT];
  __shared__ __align__(16) k1_char s_a[WARPS][MAXT];
  __shared__ __align__(16) k1_char s_b[WARPS][MAXT];
  __shared__ uint16_t s_sym[WARPS][2 * MAXT];
  __shared__ uint16_t s_rk[WARPS][2 * MAXT];
  __shared__ uint16_t s_ps[WARPS][K1_MAX_PIECES];
  __shared__ uint16_t s_pl[WARPS][K1_MAX_PIECES];
  __shared__ int32_t s_row[WARPS][K1_CTX + 3];
  __shared__ int s_meta[WARPS][4];

  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = a.n > 0 ? a.n : 1;
  const int n_cand = a.B * per;
  const int R = n_cand + (a.n > 0 ? a.B : 0);          // candidates, then (n > 0) the B unedited captions
  const int r = blockIdx.x * WARPS + w;
  if (r >= R) return;
  const bool is_base = r >= n_cand;
  const int b = is_base ? r - n_cand : r / per;
  const int off = a.cap_off[b];
  int len = a.cap_off[b + 1] - off;
  int flags = 0;
  if (len > MAXT - 24) { flags |= K1_FLAG_TOO_LONG; len = 0; }      // one inserted character and the 16-byte load tail fit
  const uint8_t* src = a.caps + off;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    for (int i = lane * 16; i < len; i += 32 * 16)        // may read up to 15 bytes past the caption: callers pad
      *reinterpret_cast<uint4*>(&s_src[w][i]) = __ldg(reinterpret_cast<const uint4*>(src + i));
  } else {
    for (int i = lane; i < len; i += 32) s_src[w][i] = __ldg(src + i);
  }
  __syncwarp();

  K1Scratch S{s_a[w], s_b[w], s_sym[w], s_rk[w], s_ps[w], s_pl[w], 0, 0};
  if (lane == 0) {
    bool edit = a.n > 0 && !is_base && (!a.valid || a.valid[r]);
    int z = 0, c = -1;
    if (a.n > 0 && !is_base) {
      z = a.sel ? a.pos[b * a.n + a.sel[b]] : a.pos[r];
      c = a.chr[r];                                        // k1_prepare checks z against the number of code points
    }
    flags |= k1_prepare(T, s_src[w], len, edit, z, c, S, a.hf_mode != 0);
    s_meta[w][0] = S.text_len;
    s_meta[w][1] = S.n_pieces;
  }
  __syncwarp();
  S.text_len = s_meta[w][0];
  S.n_pieces = s_meta[w][1];
  for (int p = lane; p < S.n_pieces; p += 32) k1_encode_piece(T, S, p);
  __syncwarp();
  if (lane == 0) {
    s_meta[w][2] = k1_emit_row(S, s_row[w]);
    if (flags && a.status_out) atomicOr(a.status_out, flags);
  }
  __syncwarp();
  int32_t* out = a.tok_out + static_cast<size_t>(r) * K1_CTX;
  for (int i = lane; i < K1_CTX; i += 32) out[i] = s_row[w][i];
  if (lane == 0) {
    a.len_out[r] = s_meta[w][2];
    if (a.base_out) a.base_out[r] = (a.n > 0 && !is_base) ? n_cand + b : -1;
  }
}

}  // namespace leaf
