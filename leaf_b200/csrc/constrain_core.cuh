// The reference's `--constrain` filter on the device: an edit candidate is valid iff it holds FEWER
// distinct dictionary words than the sentence it came from,
//     valid = len(W & set(word_tokenize(adv.lower()))) < len(W & set(word_tokenize(orig.lower())))
// (/root/reference/utils_attacks.py:110-143, applied at :321-325 and :360-364; W = set(nltk.corpus.words.words())).
//
// `nltk.word_tokenize` = Punkt sentence split + NLTKWordTokenizer (nltk/tokenize/destructive.py). The Treebank-style
// tokenizer is a fixed sequence of regular-expression substitutions; every substitution is one scalar pass below, in
// NLTK's order, over ASCII bytes. Punkt's trained parameters only exist inside NLTK's data package, so the sentence
// split is its first-pass rule with a caller-supplied abbreviation set (oracle/nltk_restate.py documents the
// approximation). PARITY UNPINNED against NLTK itself (not installable here); pinned bit-exactly against the oracle's
// `re`-based restatement by the CPU harness (tests/csrc) and on the GPU.
//
// The word list W and the abbreviations arrive as open-addressing tables of 64-bit FNV-1a hashes built on the host.
#pragma once
#include <stdint.h>

#include <vector>

#if defined(__CUDACC__)
#define CN_HD __host__ __device__ __forceinline__
#else
#define CN_HD inline
#endif

namespace leaf {

constexpr int CN_MAX_TEXT = 512;           // bytes of one (edited) sentence accepted by the filter
constexpr int CN_BUF = 2048;               // ping-pong buffers: the substitutions only ever insert spaces
constexpr int CN_MAX_FOUND = 96;           // distinct dictionary words tracked per sentence
constexpr int CN_FLAG_TOO_LONG = 8, CN_FLAG_OVERFLOW = 16;

struct CnTables {
  const uint64_t* words;                   // hash set of W (0 = empty slot)
  uint32_t words_bits;
  const uint64_t* abbrev;                  // hash set of Punkt abbreviation types (may be nullptr)
  uint32_t abbrev_bits;
};

CN_HD uint64_t cn_fnv(const uint8_t* s, int n) {
  uint64_t h = 1469598103934665603ull;
  for (int i = 0; i < n; ++i) { h ^= s[i]; h *= 1099511628211ull; }
  return h ? h : 1ull;
}
CN_HD bool cn_lookup(const uint64_t* tab, uint32_t bits, uint64_t h) {
  if (!tab) return false;
  const uint32_t mask = (1u << bits) - 1u;
  uint32_t slot = static_cast<uint32_t>(h >> 17) & mask;
  for (;;) {
#if defined(__CUDA_ARCH__)
    const uint64_t e = __ldg(reinterpret_cast<const unsigned long long*>(tab) + slot);
#else
    const uint64_t e = tab[slot];
#endif
    if (e == 0) return false;
    if (e == h) return true;
    slot = (slot + 1u) & mask;
  }
}

CN_HD bool cn_ws(uint8_t c) { return c == 32 || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); }      // str.isspace, ASCII
CN_HD bool cn_digit(uint8_t c) { return c >= '0' && c <= '9'; }
CN_HD bool cn_word(uint8_t c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || cn_digit(c) || c == '_'; }   // \w
CN_HD bool cn_nonword(uint8_t c) {          // punkt.py _re_non_word_chars
  return c == '?' || c == '!' || c == ')' || c == '"' || c == ';' || c == '}' || c == ']' || c == '*' || c == ':' || c == '@' ||
         c == '\'' || c == '(' || c == '{' || c == '[';
}
CN_HD bool cn_in(uint8_t c, const char* set) {
  for (; *set; ++set)
    if (c == static_cast<uint8_t>(*set)) return true;
  return false;
}

struct CnBuf {
  uint8_t* p;
  int n;
  bool ovf;
};
CN_HD void cn_put(CnBuf& b, uint8_t c) {
  if (b.n < CN_BUF) b.p[b.n++] = c; else b.ovf = true;
}
CN_HD void cn_pad(CnBuf& b, const uint8_t* s, int n) {      // " s "
  cn_put(b, ' ');
  for (int i = 0; i < n; ++i) cn_put(b, s[i]);
  cn_put(b, ' ');
}

// ---- single substitutions (names follow nltk/tokenize/destructive.py) -------------------------------------------------------
// re.sub(r"[set]", r" \g<0> ")
CN_HD void cn_pad_chars(const uint8_t* a, int n, CnBuf& o, const char* set) {
  for (int i = 0; i < n; ++i) {
    if (cn_in(a[i], set)) cn_pad(o, a + i, 1); else cn_put(o, a[i]);
  }
}
// re.sub(r"c{min,}", r" \g<0> ")   (STARTING_QUOTES[0] for '`' with min 1, PUNCTUATION[3] for '.' with min 2)
CN_HD void cn_pad_runs(const uint8_t* a, int n, CnBuf& o, uint8_t c, int min_run) {
  for (int i = 0; i < n;) {
    if (a[i] == c) {
      int j = i;
      while (j < n && a[j] == c) ++j;
      if (j - i >= min_run) cn_pad(o, a + i, j - i);
      else
        for (int k = i; k < j; ++k) cn_put(o, a[k]);
      i = j;
    } else {
      cn_put(o, a[i++]);
    }
  }
}
// re.sub(r"xy", " xy ") for a two-character literal, non-overlapping (STARTING_QUOTES[2] "``", DOUBLE_DASHES "--",
// ENDING_QUOTES[1] "''")
CN_HD void cn_pad_pair(const uint8_t* a, int n, CnBuf& o, uint8_t x, uint8_t y) {
  for (int i = 0; i < n;) {
    if (i + 1 < n && a[i] == x && a[i + 1] == y) { cn_pad(o, a + i, 2); i += 2; }
    else cn_put(o, a[i++]);
  }
}
// STARTING_QUOTES[1]: r'^"' -> '``'
CN_HD void cn_first_quote(const uint8_t* a, int n, CnBuf& o) {
  int i = 0;
  if (n > 0 && a[0] == '"') { cn_put(o, '`'); cn_put(o, '`'); i = 1; }
  for (; i < n; ++i) cn_put(o, a[i]);
}
// STARTING_QUOTES[3]: r"([ \(\[{<])(\"|\'{2})" -> r"\1 `` "
CN_HD void cn_open_quotes(const uint8_t* a, int n, CnBuf& o) {
  for (int i = 0; i < n;) {
    if (cn_in(a[i], " ([{<") && i + 1 < n) {
      int len = 0;
      if (a[i + 1] == '"') len = 1;
      else if (a[i + 1] == '\'' && i + 2 < n && a[i + 2] == '\'') len = 2;
      if (len) {
        cn_put(o, a[i]); cn_put(o, ' '); cn_put(o, '`'); cn_put(o, '`'); cn_put(o, ' ');
        i += 1 + len;
        continue;
      }
    }
    cn_put(o, a[i++]);
  }
}
// STARTING_QUOTES[4]: r"(?i)(\')(?!re|ve|ll|m|t|s|d|n)(\w)\b" -> r"\1 \2"
CN_HD void cn_apos_single(const uint8_t* a, int n, CnBuf& o) {
  for (int i = 0; i < n;) {
    if (a[i] == '\'' && i + 1 < n && cn_word(a[i + 1]) && !cn_in(a[i + 1], "mtsdnMTSDN") && (i + 2 == n || !cn_word(a[i + 2]))) {
      cn_put(o, '\''); cn_put(o, ' '); cn_put(o, a[i + 1]);
      i += 2;
    } else {
      cn_put(o, a[i++]);
    }
  }
}
// PUNCTUATION[0] (with_space = true):  r'([^\.])(\.)([\]\)}>"\' ]*)\s*$' -> r"\1 \2 \3 "
// PUNCTUATION[5] (with_space = false): r'([^\.])(\.)([\]\)}>"\']*)\s*$'  -> r"\1 \2\3 "
// Only the LAST period of the text can match (an earlier one would have a period in its tail).
CN_HD void cn_final_period(const uint8_t* a, int n, CnBuf& o, bool with_space) {
  int d = n - 1;
  while (d >= 0 && a[d] != '.') --d;
  bool hit = d >= 1 && a[d - 1] != '.';
  int e = d + 1;
  if (hit) {
    while (e < n && (cn_in(a[e], "])}>\"'") || (with_space && a[e] == ' '))) ++e;
    int f = e;
    while (f < n && cn_ws(a[f])) ++f;
    hit = f == n;
  }
  if (!hit) {
    for (int i = 0; i < n; ++i) cn_put(o, a[i]);
    return;
  }
  for (int i = 0; i < d; ++i) cn_put(o, a[i]);
  cn_put(o, ' '); cn_put(o, '.');
  if (with_space) cn_put(o, ' ');
  for (int i = d + 1; i < e; ++i) cn_put(o, a[i]);
  cn_put(o, ' ');
}
// PUNCTUATION[1]: r"([:,])([^\d])" -> r" \1 \2"
CN_HD void cn_colon_comma(const uint8_t* a, int n, CnBuf& o) {
  for (int i = 0; i < n;) {
    if ((a[i] == ':' || a[i] == ',') && i + 1 < n && !cn_digit(a[i + 1])) {
      cn_put(o, ' '); cn_put(o, a[i]); cn_put(o, ' '); cn_put(o, a[i + 1]);
      i += 2;
    } else {
      cn_put(o, a[i++]);
    }
  }
}
// PUNCTUATION[2]: r"([:,])$" -> r" \1 "
CN_HD void cn_colon_comma_end(const uint8_t* a, int n, CnBuf& o) {
  const bool hit = n > 0 && (a[n - 1] == ':' || a[n - 1] == ',');
  for (int i = 0; i < n - (hit ? 1 : 0); ++i) cn_put(o, a[i]);
  if (hit) cn_pad(o, a + n - 1, 1);
}
// PUNCTUATION[7]: r"([^'])' " -> r"\1 ' "
CN_HD void cn_apos_space(const uint8_t* a, int n, CnBuf& o) {
  for (int i = 0; i < n;) {
    if (i + 2 < n && a[i] != '\'' && a[i + 1] == '\'' && a[i + 2] == ' ') {
      cn_put(o, a[i]); cn_put(o, ' '); cn_put(o, '\''); cn_put(o, ' ');
      i += 3;
    } else {
      cn_put(o, a[i++]);
    }
  }
}
// ENDING_QUOTES[2]: r'"' -> " '' "
CN_HD void cn_dquote(const uint8_t* a, int n, CnBuf& o) {
  for (int i = 0; i < n; ++i) {
    if (a[i] == '"') { cn_put(o, ' '); cn_put(o, '\''); cn_put(o, '\''); cn_put(o, ' '); }
    else cn_put(o, a[i]);
  }
}
// ENDING_QUOTES[3]: r"([^' ])('[sS]|'[mM]|'[dD]|') " -> r"\1 \2 "
CN_HD void cn_possessive(const uint8_t* a, int n, CnBuf& o) {
  for (int i = 0; i < n;) {
    if (a[i] != '\'' && a[i] != ' ' && i + 1 < n && a[i + 1] == '\'') {
      if (i + 3 < n && cn_in(a[i + 2], "smdSMD") && a[i + 3] == ' ') {
        cn_put(o, a[i]); cn_put(o, ' '); cn_put(o, '\''); cn_put(o, a[i + 2]); cn_put(o, ' ');
        i += 4;
        continue;
      }
      if (i + 2 < n && a[i + 2] == ' ') {
        cn_put(o, a[i]); cn_put(o, ' '); cn_put(o, '\''); cn_put(o, ' ');
        i += 3;
        continue;
      }
    }
    cn_put(o, a[i++]);
  }
}
CN_HD bool cn_match(const uint8_t* a, int n, int i, const char* lit) {       // case-insensitive literal at a[i..]
  for (int k = 0; lit[k]; ++k) {
    if (i + k >= n) return false;
    uint8_t c = a[i + k];
    if (c >= 'A' && c <= 'Z') c += 32;
    if (c != static_cast<uint8_t>(lit[k])) return false;
  }
  return true;
}
CN_HD int cn_len(const char* s) { int n = 0; while (s[n]) ++n; return n; }
// ENDING_QUOTES[4]: r"([^' ])('ll|'LL|'re|'RE|'ve|'VE|n't|N'T) " -> r"\1 \2 "
CN_HD void cn_contraction_suffix(const uint8_t* a, int n, CnBuf& o) {
  for (int i = 0; i < n;) {
    if (a[i] != '\'' && a[i] != ' ' && i + 4 < n && a[i + 4] == ' ' &&
        (cn_match(a, n, i + 1, "'ll") || cn_match(a, n, i + 1, "'re") || cn_match(a, n, i + 1, "'ve") || cn_match(a, n, i + 1, "n't"))) {
      // NLTK lists only the all-lower and all-upper spellings; lower-cased input never holds the mixed ones
      cn_put(o, a[i]); cn_put(o, ' '); cn_put(o, a[i + 1]); cn_put(o, a[i + 2]); cn_put(o, a[i + 3]); cn_put(o, ' ');
      i += 5;
    } else {
      cn_put(o, a[i++]);
    }
  }
}
// CONTRACTIONS2: r"(?i)\b(x)(?#X)(y)\b" -> r" \1 \2 "   (ws_after: the trailing \b is (?=\s) instead, "wanna")
CN_HD void cn_contraction2(const uint8_t* a, int n, CnBuf& o, const char* x, const char* y, bool ws_after) {
  const int lx = cn_len(x), ly = cn_len(y);
  for (int i = 0; i < n;) {
    const bool b0 = i == 0 || !cn_word(a[i - 1]);                       // \b before a word character
    if (b0 && cn_match(a, n, i, x) && cn_match(a, n, i + lx, y)) {
      const int e = i + lx + ly;
      const bool b1 = ws_after ? (e < n && cn_ws(a[e])) : (e == n || !cn_word(a[e]));
      if (b1) {
        cn_pad(o, a + i, lx);
        for (int k = 0; k < ly; ++k) cn_put(o, a[i + lx + k]);
        cn_put(o, ' ');
        i = e;
        continue;
      }
    }
    cn_put(o, a[i++]);
  }
}
// CONTRACTIONS3: r"(?i) ('t)(?#X)(is|was)\b" -> r" \1 \2 "
CN_HD void cn_contraction3(const uint8_t* a, int n, CnBuf& o, const char* y) {
  const int ly = cn_len(y);
  for (int i = 0; i < n;) {
    if (a[i] == ' ' && cn_match(a, n, i + 1, "'t") && cn_match(a, n, i + 3, y) && (i + 3 + ly == n || !cn_word(a[i + 3 + ly]))) {
      cn_put(o, ' '); cn_put(o, a[i + 1]); cn_put(o, a[i + 2]); cn_put(o, ' ');
      for (int k = 0; k < ly; ++k) cn_put(o, a[i + 3 + k]);
      cn_put(o, ' ');
      i += 3 + ly;
    } else {
      cn_put(o, a[i++]);
    }
  }
}

// Distinct dictionary words seen so far
struct CnFound {
  uint64_t h[CN_MAX_FOUND];
  int n;
  bool ovf;
};
CN_HD void cn_add_found(CnFound& f, uint64_t h) {
  for (int i = 0; i < f.n; ++i)
    if (f.h[i] == h) return;
  if (f.n < CN_MAX_FOUND) f.h[f.n++] = h; else f.ovf = true;
}

// NLTKWordTokenizer.tokenize(sentence).  buf_a holds the sentence (n bytes); buf_b is the other ping-pong buffer.
// Dictionary tokens are added to `found`. Returns true on buffer overflow.
CN_HD bool cn_treebank_count(const CnTables& T, uint8_t* buf_a, int n, uint8_t* buf_b, CnFound& found) {
  uint8_t* cur = buf_a;
  uint8_t* nxt = buf_b;
  bool ovf = false;
#define CN_PASS(call)                                   \
  {                                                     \
    CnBuf o{nxt, 0, false};                             \
    call;                                               \
    ovf |= o.ovf;                                       \
    uint8_t* t_ = cur; cur = nxt; nxt = t_; n = o.n;    \
  }
  // STARTING_QUOTES
  CN_PASS(cn_pad_runs(cur, n, o, '`', 1))
  CN_PASS(cn_first_quote(cur, n, o))
  CN_PASS(cn_pad_pair(cur, n, o, '`', '`'))
  CN_PASS(cn_open_quotes(cur, n, o))
  CN_PASS(cn_apos_single(cur, n, o))
  // PUNCTUATION
  CN_PASS(cn_final_period(cur, n, o, true))
  CN_PASS(cn_colon_comma(cur, n, o))
  CN_PASS(cn_colon_comma_end(cur, n, o))
  CN_PASS(cn_pad_runs(cur, n, o, '.', 2))
  CN_PASS(cn_pad_chars(cur, n, o, ";@#$%&"))
  CN_PASS(cn_final_period(cur, n, o, false))
  CN_PASS(cn_pad_chars(cur, n, o, "?!"))
  CN_PASS(cn_apos_space(cur, n, o))
  CN_PASS(cn_pad_chars(cur, n, o, "*"))
  // PARENS_BRACKETS, DOUBLE_DASHES
  CN_PASS(cn_pad_chars(cur, n, o, "][(){}<>"))
  CN_PASS(cn_pad_pair(cur, n, o, '-', '-'))
  // " " + text + " "
  CN_PASS({ cn_pad(o, cur, n); })
  // ENDING_QUOTES
  CN_PASS(cn_pad_pair(cur, n, o, '\'', '\''))
  CN_PASS(cn_dquote(cur, n, o))
  CN_PASS(cn_possessive(cur, n, o))
  CN_PASS(cn_contraction_suffix(cur, n, o))
  // CONTRACTIONS2, CONTRACTIONS3
  CN_PASS(cn_contraction2(cur, n, o, "can", "not", false))
  CN_PASS(cn_contraction2(cur, n, o, "d", "'ye", false))
  CN_PASS(cn_contraction2(cur, n, o, "gim", "me", false))
  CN_PASS(cn_contraction2(cur, n, o, "gon", "na", false))
  CN_PASS(cn_contraction2(cur, n, o, "got", "ta", false))
  CN_PASS(cn_contraction2(cur, n, o, "lem", "me", false))
  CN_PASS(cn_contraction2(cur, n, o, "more", "'n", false))
  CN_PASS(cn_contraction2(cur, n, o, "wan", "na", true))
  CN_PASS(cn_contraction3(cur, n, o, "is"))
  CN_PASS(cn_contraction3(cur, n, o, "was"))
#undef CN_PASS
  // text.split(): tokens; only those in W count, and W holds plain words
  for (int i = 0; i < n;) {
    while (i < n && cn_ws(cur[i])) ++i;
    int j = i;
    while (j < n && !cn_ws(cur[j])) ++j;
    if (j > i) {
      const uint64_t h = cn_fnv(cur + i, j - i);
      if (cn_lookup(T.words, T.words_bits, h)) cn_add_found(found, h);
    }
    i = j;
  }
  return ovf || found.ovf;
}

// `-?[\.,]?\d[\d,\.-]*` (Punkt's numeric token type without the final period)
CN_HD bool cn_is_number(const uint8_t* s, int n) {
  int i = 0;
  if (i < n && s[i] == '-') ++i;
  if (i < n && (s[i] == '.' || s[i] == ',')) ++i;
  if (i >= n || !cn_digit(s[i])) return false;
  for (++i; i < n; ++i)
    if (!(cn_digit(s[i]) || s[i] == ',' || s[i] == '.' || s[i] == '-')) return false;
  return true;
}

// count = len(W & set(word_tokenize(text)))  for lower-cased `text` (n bytes, in buf_text; clobbered). Mirrors
// oracle/nltk_restate.py::sent_split + treebank_tokenize. Returns the count; flags gets CN_FLAG_* on trouble.
CN_HD int cn_count_words(const CnTables& T, const uint8_t* text, int n, uint8_t* buf_a, uint8_t* buf_b, int& flags) {
  CnFound found;
  found.n = 0;
  found.ovf = false;
  bool ovf = false;
  int last = 0;
  auto sentence = [&](int s, int e) {
    if (e <= s) return;
    for (int k = s; k < e; ++k) buf_a[k - s] = text[k];
    ovf |= cn_treebank_count(T, buf_a, e - s, buf_b, found);
  };
  // punkt.py _period_context_fmt: a sentence-end character followed by a non-word character, or by whitespace and a token
  auto end_context = [&](int i) {
    const uint8_t c = text[i];
    if (!(c == '.' || c == '?' || c == '!') || i + 1 >= n) return false;
    if (cn_nonword(text[i + 1])) return true;
    int j = i + 1;
    while (j < n && cn_ws(text[j])) ++j;
    return j > i + 1 && j < n;
  };
  bool pending = false;      // a break decided at an earlier end character of the same whitespace-delimited word
  for (int i = 0; i < n;) {
    const uint8_t c = text[i];
    if (end_context(i)) {
      const uint8_t nx = text[i + 1];
      int j = i + 1;
      while (j < n && cn_ws(text[j])) ++j;
      const bool after_ws = j > i + 1 && j < n;
      bool brk = true;
      if (c == '.') {
        if ((i > 0 && text[i - 1] == '.') || nx == '.') {
          brk = false;
        } else {
          int s = i;
          while (s > 0 && !cn_ws(text[s - 1]) && !cn_nonword(text[s - 1])) --s;
          const int sl = i - s;
          if (sl > 0) {
            int hy = i;                                                  // the part after the last '-'
            while (hy > s && text[hy - 1] != '-') --hy;
            if (cn_lookup(T.abbrev, T.abbrev_bits, cn_fnv(text + s, sl)) ||
                (hy > s && cn_lookup(T.abbrev, T.abbrev_bits, cn_fnv(text + hy, i - hy))))
              brk = false;
            else if (sl == 1 && ((text[s] >= 'a' && text[s] <= 'z') || (text[s] >= 'A' && text[s] <= 'Z')))
              brk = false;
            else if (cn_is_number(text + s, sl))
              brk = false;
          }
        }
      }
      // potential ends inside ONE word are one decision, taken at the last of them (NLTK >= 3.6.6,
      // PunktSentenceTokenizer._match_potential_end_contexts; oracle/nltk_restate.py::sent_split)
      bool later = false;
      for (int k = i + 1; k < n && !cn_ws(text[k]); ++k)
        if (end_context(k)) { later = true; break; }
      if (later) {
        pending |= brk;
        ++i;
        continue;
      }
      brk |= pending;
      pending = false;
      if (brk) {
        int end = i + 1;
        int start = after_ws ? j : i + 1;
        int k = start;
        while (k < n && cn_in(text[k], "\"')]}")) ++k;
        if (k > start && (k == n || cn_ws(text[k]) || (k + 1 < n && text[k] == '-' && text[k + 1] == '-'))) {
          end = k;
          while (k < n && cn_ws(text[k])) ++k;
          start = k;
        }
        sentence(last, end);
        last = start;
        i = start > i + 1 ? start : i + 1;
        continue;
      }
    }
    ++i;
  }
  int e = n;
  while (e > last && cn_ws(text[e - 1])) --e;
  sentence(last, e);
  if (ovf) flags |= CN_FLAG_OVERFLOW;
  return found.n;
}

// host side: the open-addressing hash set cn_lookup probes (word blobs back to back + [n+1] offsets)
inline std::vector<uint64_t> cn_build_table(const uint8_t* blob, const int32_t* off, int n, uint32_t* bits_out) {
  uint32_t bits = 4;
  while ((1ull << bits) < static_cast<uint64_t>(n) * 2 + 16) ++bits;
  std::vector<uint64_t> tab(1ull << bits, 0ull);
  const uint32_t mask = (1u << bits) - 1u;
  for (int i = 0; i < n; ++i) {
    const uint64_t h = cn_fnv(blob + off[i], off[i + 1] - off[i]);
    uint32_t slot = static_cast<uint32_t>(h >> 17) & mask;
    while (tab[slot] != 0 && tab[slot] != h) slot = (slot + 1u) & mask;
    tab[slot] = h;
  }
  *bits_out = bits;
  return tab;
}

}  // namespace leaf
