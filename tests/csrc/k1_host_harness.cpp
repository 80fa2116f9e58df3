// TEST INFRASTRUCTURE ONLY - compiles the K1 scalar core (leaf_b200/csrc/k1_core.cuh) for the CPU so that the
// `-m "not gpu"` suite can pin the integer logic of the tokenization kernel against the oracle and the
// golden vectors generated from the reference. Never loaded by the product (leaf_b200/ has no CPU path).
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../leaf_b200/csrc/constrain_core.cuh"
#include "../../leaf_b200/csrc/k1_tables_host.h"

using namespace leaf;

static std::vector<uint64_t> g_tab;
static bool g_hf = false;

extern "C" int k1h_set_mode(int hf) { g_hf = hf != 0; return 0; }

extern "C" int k1h_load(const uint32_t* merge_pairs, int n) {
  g_tab = k1_build_merge_table(merge_pairs, n);
  return 0;
}

// same contract as leaf_expand_tokenize, host pointers; sequential emulation of "warp per candidate"
extern "C" int k1h_expand_tokenize(const uint8_t* caps, const int32_t* cap_off, int B, int n, const int32_t* pos,
                                   const int32_t* chr, const int32_t* sel, const uint8_t* valid, int32_t* tok_out,
                                   int32_t* len_out) {
  K1Tables T = k1_host_tables(g_tab.data());
  constexpr int MAXT = 3584;                 // the long-text variant's buffers (k1_tokenize.cuh: K1_LONG_TEXT)
  std::vector<k1_char> a(MAXT), b(MAXT);
  std::vector<uint16_t> sym(2 * MAXT), rk(2 * MAXT), ps(K1_MAX_PIECES), pl(K1_MAX_PIECES);
  K1Scratch S{a.data(), b.data(), sym.data(), rk.data(), ps.data(), pl.data(), 0, 0};
  int flags = 0;
  const int per = n > 0 ? n : 1;
  for (int r = 0; r < B * per; ++r) {
    const int bb = r / per, j = r % per;
    const uint8_t* src = caps + cap_off[bb];
    int len = cap_off[bb + 1] - cap_off[bb];
    if (len > MAXT - 24) { flags |= K1_FLAG_TOO_LONG; len = 0; }
    bool edit = n > 0 && (!valid || valid[r]);
    int z = 0, c = -1;
    if (n > 0) {
      z = sel ? pos[bb * n + sel[bb]] : pos[r];
      c = chr[r];
    }
    flags |= k1_prepare(T, src, len, edit, z, c, S, g_hf);
    for (int p = 0; p < S.n_pieces; ++p) k1_encode_piece(T, S, p);
    len_out[r] = k1_emit_row(S, tok_out + (size_t)r * K1_CTX);
  }
  return flags;
}

// ---- the --constrain filter core (leaf_b200/csrc/constrain_core.cuh) on the CPU --------------------------------------------
static std::vector<uint64_t> g_words, g_abbrev;
static CnTables g_cn{};

extern "C" int cnh_load(const uint8_t* wb, const int32_t* wo, int nw, const uint8_t* ab, const int32_t* ao, int na) {
  g_words = cn_build_table(wb, wo, nw, &g_cn.words_bits);
  g_cn.words = g_words.data();
  g_cn.abbrev = nullptr;
  if (na > 0) {
    g_abbrev = cn_build_table(ab, ao, na, &g_cn.abbrev_bits);
    g_cn.abbrev = g_abbrev.data();
  }
  return 0;
}

// same contract as leaf_constrain_mask's count output, host pointers; sequential emulation of "thread per sentence"
extern "C" int cnh_counts(const uint8_t* caps, const int32_t* cap_off, int B, int n, const int32_t* pos, const int32_t* chr,
                          const int32_t* sel, int32_t* count_out) {
  std::vector<uint8_t> text(CN_MAX_TEXT + 8), a(CN_BUF), b(CN_BUF);
  int flags = 0;
  for (int r = 0; r < B * n + B; ++r) {
    const bool is_base = r >= B * n;
    const int bb = is_base ? r - B * n : r / n;
    const uint8_t* src = caps + cap_off[bb];
    int len = cap_off[bb + 1] - cap_off[bb];
    if (len > CN_MAX_TEXT - 1) { flags |= CN_FLAG_TOO_LONG; len = 0; }
    int m;
    if (is_base) {
      memcpy(text.data(), src, len);
      m = len;
    } else {
      const int z = sel ? pos[bb * n + sel[bb]] : pos[r];
      m = k1_apply_edit(src, len, z, chr[r], text.data());
    }
    for (int i = 0; i < m; ++i)
      if (text[i] >= 'A' && text[i] <= 'Z') text[i] += 32;
    count_out[r] = cn_count_words(g_cn, text.data(), m, a.data(), b.data(), flags);
  }
  return flags;
}
