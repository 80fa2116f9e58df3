// K1 core: one LEAF edit candidate -> CLIP BPE token row. Scalar building blocks shared by the CUDA kernel
// (k1_tokenize.cu: one warp per candidate, lanes over regex pieces) and by the CPU-compiled test harness
// (tests/csrc/k1_host_harness.cpp), so that the integer logic can be pinned against the oracle without a GPU.
//
// Follows, step by step:
//   edit rule          /root/reference/utils_attacks.py:169-213 (generate_sentence, alternative = -1)
//   basic_clean        /root/reference/src/open_clip/tokenizer.py:66-69  (ftfy = identity on the domain below,
//                      html.unescape twice, strip)
//   whitespace_clean   tokenizer.py:72-75      lower: tokenizer.py:83-85
//   regex split        tokenizer.py:160-163    bpe: tokenizer.py:172-211     row layout: tokenizer.py:256-263
// html.unescape is CPython's Lib/html/__init__.py (_charref regex + _replace_charref) restated on code points
// <= U+024F. Domain: UTF-8 captions whose code points are all <= U+024F (ASCII, Latin-1 Supplement, Latin Extended-A / -B:
// Western and Central European text), one LEAF edit, and whatever the two unescape passes make of that while staying
// <= U+024F. Text is held as 16-bit code points internally. Everything outside raises a status flag (the host raises
// LeafError), never a silently different row: code points > U+024F / malformed UTF-8 / the 24 capitals whose str.lower()
// leaves the range (U+0130 and the like); and the inputs on which ftfy.fix_text - which the
// reference runs first and which cannot be restated here - is NOT the identity: C1 controls U+0080..U+009F (ftfy maps
// them to Windows-1252), a UTF-8-lead-like character followed by a continuation-like one (`Ã©`, and the Windows-1252
// spellings of continuation bytes such as `Å¡` / `Ãœ`: ftfy re-decodes mojibake), a third level of entity nesting and ALL-CAPS entity names (`&EACUTE;`) in text without '<' (ftfy unescapes
// one level itself, with upper-case variants html.unescape does not know).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define K1_HD __host__ __device__ __forceinline__
#else
#define K1_HD inline
#endif

namespace leaf {

constexpr int K1_MAX_TEXT = 1024;          // bytes per candidate buffer (caption <= 1000 bytes + 1 inserted)
constexpr int K1_MAX_PIECES = 76;          // only the first 75 ids of a row can survive truncation
constexpr int K1_CTX = 77;
constexpr int K1_SOT = 49406, K1_EOT = 49407;
constexpr uint16_t K1_UNSUP_V = 0xFFFF, K1_EMPTY_V = 0xFFFE;
constexpr int K1_MAX_CP = 0x24F;            // largest code point of the domain; tables have K1_MAX_CP + 1 entries
typedef uint16_t k1_char;                   // one code point of the text being tokenized
constexpr uint16_t K1_NO_RANK = 0xFFFF, K1_DIRTY = 0xFFFE;
constexpr int K1_FLAG_ENTITY_DOMAIN = 1, K1_FLAG_NON_ASCII = 2 /* outside U+0000..U+00FF, or ftfy would rewrite it */, K1_FLAG_TOO_LONG = 4;
constexpr uint64_t K1_SLOT_EMPTY = ~0ull;

struct K1Tables {
  const uint16_t* byte_id;    // [256] id of each byte symbol (bytes_to_unicode order, tokenizer.py:31-51)
  const uint8_t* cls;         // [592] regex class of the code point: 0 other, 1 \p{L}, 2 \p{N}, 3 \s
  const uint8_t* ws;          // [592] str.split()/strip() whitespace
  const uint16_t* lower;      // [592] str.lower(), K1_UNSUP_V where it leaves the domain
  const uint16_t* numref;     // [256] outcome of &#N; for N < 256
  const uint16_t* ent_off;    // [n_ent] offsets into ent_blob (names sorted bytewise)
  const uint8_t* ent_len;     // [n_ent]
  const uint16_t* ent_val;    // [n_ent] code point, or K1_UNSUP_V
  const uint8_t* ent_blob;
  int n_ent;
  const uint64_t* merge_tab;  // open-addressing table: (key << 32) | rank, key = left << 16 | right
  uint32_t merge_bits;        // log2(slots)
};

// per-candidate scratch (shared memory on the device)
struct K1Scratch {
  k1_char* buf_a;             // [K1_MAX_TEXT]
  k1_char* buf_b;             // [K1_MAX_TEXT]
  uint16_t* sym;              // [2 * K1_MAX_TEXT] symbols of piece p live at sym + 2 * start(p)
  uint16_t* rk;               // [2 * K1_MAX_TEXT] rank cache, same indexing
  uint16_t* piece_start;      // [K1_MAX_PIECES]
  uint16_t* piece_len;        // [K1_MAX_PIECES]  on return of k1_encode_piece: number of ids (special: 0x8000|which)
  int text_len;
  int n_pieces;
};

K1_HD uint32_t k1_hash(uint32_t key, uint32_t bits) { return (key * 2654435761u) >> (32u - bits); }

K1_HD uint32_t k1_merge_rank(const K1Tables& T, uint32_t left, uint32_t right) {
  const uint32_t key = (left << 16) | right;
  const uint32_t mask = (1u << T.merge_bits) - 1u;
  uint32_t slot = k1_hash(key, T.merge_bits);
  for (;;) {
#if defined(__CUDA_ARCH__)
    const uint64_t e = __ldg(reinterpret_cast<const unsigned long long*>(T.merge_tab) + slot);
#else
    const uint64_t e = T.merge_tab[slot];
#endif
    if (e == K1_SLOT_EMPTY) return K1_NO_RANK;
    if (static_cast<uint32_t>(e >> 32) == key) return static_cast<uint32_t>(e & 0xFFFFu);
    slot = (slot + 1u) & mask;
  }
}

// ---- UTF-8 -> 16-bit code points (U+0000..U+024F) ---------------------------------------------------------
K1_HD int k1_decode_utf8(const uint8_t* in, int n, k1_char* out, int* flags) {
  int m = 0;
  for (int i = 0; i < n;) {
    const uint8_t b = in[i];
    if (b < 0x80) { out[m++] = b; ++i; }
    else if (b >= 0xC2 && b <= 0xC9 && i + 1 < n && (in[i + 1] & 0xC0) == 0x80 &&
             (((b & 0x1F) << 6) | (in[i + 1] & 0x3F)) <= K1_MAX_CP) {
      out[m++] = static_cast<k1_char>(((b & 0x1F) << 6) | (in[i + 1] & 0x3F));
      i += 2;
    } else {                                                // > U+024F or malformed: flagged, one '?' per sequence
      *flags |= K1_FLAG_NON_ASCII;
      out[m++] = '?';
      ++i;
      while (i < n && (in[i] & 0xC0) == 0x80) ++i;
    }
  }
  return m;
}

// ---- edit rule (utils_attacks.py:169-213) ---------------------------------------------
// z even = slot before character z/2, z odd = character z/2. c = code point or -1.
template <typename Ch>
K1_HD int k1_apply_edit(const Ch* S, int len, int z, int c, Ch* out) {
  const int i = z >> 1;
  int n = 0;
  if (z & 1) {
    const bool del = (c == -1) || (c == static_cast<int>(S[i]));
    for (int j = 0; j < i; ++j) out[n++] = S[j];
    if (!del) out[n++] = static_cast<Ch>(c);
    for (int j = i + 1; j < len; ++j) out[n++] = S[j];
  } else {
    const bool noop = (c == -1) || (c == '_');
    for (int j = 0; j < i; ++j) out[n++] = S[j];
    if (!noop) out[n++] = static_cast<Ch>(c);
    for (int j = i; j < len; ++j) out[n++] = S[j];
  }
  return n;
}

// ---- html.unescape, one pass -----------------------------------------------------------------------------
K1_HD int k1_find_entity(const K1Tables& T, const k1_char* s, int n) {
  int lo = 0, hi = T.n_ent - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const uint8_t* e = T.ent_blob + T.ent_off[mid];
    const int el = T.ent_len[mid];
    int cmp = 0;
    const int m = n < el ? n : el;
    for (int k = 0; k < m; ++k) {
      if (s[k] != e[k]) { cmp = s[k] < e[k] ? -1 : 1; break; }      // names are ASCII: a wider code point sorts after them
    }
    if (cmp == 0) cmp = (n < el) ? -1 : (n > el ? 1 : 0);
    if (cmp == 0) return mid;
    if (cmp < 0) hi = mid - 1; else lo = mid + 1;
  }
  return -1;
}

K1_HD bool k1_is_digit(k1_char c) { return c >= '0' && c <= '9'; }
K1_HD int k1_hexval(k1_char c) {
  if (c >= '0' && c <= '9') return c - '0';
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  if (c >= 'A' && c <= 'F') return c - 'A' + 10;
  return -1;
}

K1_HD int k1_emit_cp(uint16_t v, k1_char* out, int m, int* flags) {
  if (v == K1_EMPTY_V) return m;
  if (v == K1_UNSUP_V) { *flags |= K1_FLAG_ENTITY_DOMAIN; out[m++] = '?'; return m; }
  out[m++] = v;
  return m;
}

K1_HD int k1_unescape(const K1Tables& T, const k1_char* in, int n, k1_char* out, int* flags, bool ftfy_caps = false) {
  int m = 0, i = 0;
  while (i < n) {
    const k1_char ch = in[i];
    if (ch != '&' || i + 1 >= n) { out[m++] = ch; ++i; continue; }
    if (in[i + 1] == '#') {
      int j = i + 2;
      uint32_t num = 0;
      bool ok = false;
      if (j < n && k1_is_digit(in[j])) {                       // &#[0-9]+;?
        while (j < n && k1_is_digit(in[j])) {
          num = num * 10u + static_cast<uint32_t>(in[j] - '0');
          if (num > 0x110000u) num = 0x110000u;
          ++j;
        }
        ok = true;
      } else if (j + 1 < n && (in[j] == 'x' || in[j] == 'X') && k1_hexval(in[j + 1]) >= 0) {   // &#[xX][0-9a-fA-F]+;?
        ++j;
        while (j < n && k1_hexval(in[j]) >= 0) {
          num = num * 16u + static_cast<uint32_t>(k1_hexval(in[j]));
          if (num > 0x110000u) num = 0x110000u;
          ++j;
        }
        ok = true;
      }
      if (!ok) { out[m++] = ch; ++i; continue; }
      if (j < n && in[j] == ';') ++j;
      uint16_t v;
      if (num < 256u) v = T.numref[num];
      else if (num > 0x10FFFFu) v = K1_UNSUP_V;                                  // U+FFFD
      else if ((num >= 0xFDD0u && num <= 0xFDEFu) || (num & 0xFFFEu) == 0xFFFEu) v = K1_EMPTY_V;
      else if (num <= static_cast<uint32_t>(K1_MAX_CP)) v = static_cast<uint16_t>(num);   // chr(num), inside the domain
      else v = K1_UNSUP_V;                                                       // chr(num) > U+024F (or U+FFFD)
      m = k1_emit_cp(v, out, m, flags);
      i = j;
      continue;
    }
    // named: [^\t\n\f <&#;]{1,32};?
    int j = i + 1, cnt = 0;
    while (j < n && cnt < 32) {
      const k1_char c = in[j];
      if (c == '\t' || c == '\n' || c == '\f' || c == ' ' || c == '<' || c == '&' || c == '#' || c == ';') break;
      ++j; ++cnt;
    }
    if (cnt == 0) { out[m++] = ch; ++i; continue; }
    if (j < n && in[j] == ';') ++j;
    const k1_char* s = in + i + 1;
    const int sl = j - (i + 1);
    int e = k1_find_entity(T, s, sl);
    if (e >= 0) {
      m = k1_emit_cp(T.ent_val[e], out, m, flags);
    } else {
      if (ftfy_caps && sl >= 3 && s[sl - 1] == ';') {         // &EACUTE; - ftfy knows upper-case variants, html.unescape does not
        bool caps = true, letter = false;
        for (int k = 0; k + 1 < sl; ++k) {
          const k1_char c = s[k];
          if (c >= 'A' && c <= 'Z') letter = true;
          else if (!(c >= '0' && c <= '9')) { caps = false; break; }
        }
        if (caps && letter) *flags |= K1_FLAG_ENTITY_DOMAIN;
      }
      int x = sl - 1;
      for (; x > 1; --x) {
        e = k1_find_entity(T, s, x);
        if (e >= 0) break;
      }
      if (x > 1) {
        m = k1_emit_cp(T.ent_val[e], out, m, flags);
        for (int k = x; k < sl; ++k) out[m++] = s[k];
      } else {
        out[m++] = '&';
        for (int k = 0; k < sl; ++k) out[m++] = s[k];
      }
    }
    i = j;
  }
  return m;
}

// ---- strip + " ".join(split()) + lower -------------------------------------------------------------------
K1_HD int k1_clean(const K1Tables& T, const k1_char* in, int n, k1_char* out, int* flags) {
  int m = 0;
  bool pending_space = false;
  for (int i = 0; i < n; ++i) {
    const k1_char c = in[i];
    if (T.ws[c]) { pending_space = (m > 0); continue; }
    if (pending_space) { out[m++] = ' '; pending_space = false; }
    k1_char lo = T.lower[c];
    if (lo == K1_UNSUP_V) { *flags |= K1_FLAG_NON_ASCII; lo = '?'; }      // str.lower() leaves the domain (U+0130 -> i + U+0307, ...)
    out[m++] = lo;
  }
  return m;
}

// The reference's pattern is compiled with IGNORECASE and applied to lower-cased text: inside the domain the one character that
// still matches an ASCII letter of the pattern's literals is U+017F (long s, folds to 's'): `'\u017f` is the contraction 's and
// `<\u017ftart_of_text>` the special token (tools/make_tables.py asserts there is no other).
K1_HD k1_char k1_fold(k1_char c) { return c == 0x17F ? static_cast<k1_char>('s') : c; }
K1_HD bool k1_match(const k1_char* t, int n, int pos, const char* lit, int ll) {
  if (pos + ll > n) return false;
  for (int k = 0; k < ll; ++k)
    if (k1_fold(t[pos + k]) != static_cast<k1_char>(lit[k])) return false;
  return true;
}

// ---- regex split of the cleaned text (tokenizer.py:160-163), first K1_MAX_PIECES pieces ------------------
// piece_len gets 0x8000 | 0 for <start_of_text>, 0x8000 | 1 for <end_of_text> (ids straight from the cache,
// tokenizer.py:159).
K1_HD int k1_split(const K1Tables& T, const k1_char* t, int n, uint16_t* piece_start, uint16_t* piece_len, bool hf = false) {
  int np = 0, pos = 0;
  while (pos < n && np < K1_MAX_PIECES) {
    const k1_char c = t[pos];
    const int cl = T.cls[c];
    if (cl == 3) { ++pos; continue; }
    int len = 0;
    uint16_t special = 0;
    // transformers' CLIPTokenizer spells the two special tokens <|startoftext|> / <|endoftext|> (same lengths)
    if (c == '<' && k1_match(t, n, pos, hf ? "<|startoftext|>" : "<start_of_text>", 15)) { len = 15; special = 0x8000; }
    else if (c == '<' && k1_match(t, n, pos, hf ? "<|endoftext|>" : "<end_of_text>", 13)) { len = 13; special = 0x8001; }
    if (special) {                     // the PIECE is cut case-insensitively, the id shortcut (tokenizer.py:159: the cache seeded
      for (int k = 0; k < len; ++k)    // with the two literals) needs the exact spelling: `<\u017ftart_of_text>` is one ordinary piece
        if (t[pos + k] == 0x17F) special = 0;
    }
    else if (c == '\'' && pos + 1 < n) {
      const k1_char d = k1_fold(t[pos + 1]);
      if (d == 's' || d == 't' || d == 'm' || d == 'd') len = 2;
      else if (pos + 2 < n) {
        const k1_char e = t[pos + 2];
        if ((d == 'r' && e == 'e') || (d == 'v' && e == 'e') || (d == 'l' && e == 'l')) len = 3;
      }
    }
    if (len == 0) {
      if (cl == 1) { len = 1; while (pos + len < n && T.cls[t[pos + len]] == 1) ++len; }
      else if (cl == 2) { len = 1; }
      else { len = 1; while (pos + len < n && T.cls[t[pos + len]] == 0) ++len; }
    }
    piece_start[np] = static_cast<uint16_t>(pos);
    piece_len[np] = special ? special : static_cast<uint16_t>(len);
    ++np;
    pos += len;
  }
  return np;
}

// what a UTF-8 continuation byte looks like after a wrong Latin-1 / Windows-1252 decode (inside the domain): U+0080..U+00BF, and
// the Windows-1252 characters of the bytes 0x8C 0x9C 0x8A 0x9A 0x9F 0x8E 0x9E 0x83
K1_HD bool k1_continuation_like(k1_char c) {
  return (c >= 0x80 && c <= 0xBF) || c == 0x152 || c == 0x153 || c == 0x160 || c == 0x161 || c == 0x178 || c == 0x17D || c == 0x17E ||
         c == 0x192;
}

// ---- everything a candidate needs before BPE (serial; lane 0 on the device) ------------------------------
// src = caption bytes; do_edit selects the LEAF edit (z, c); returns flags.
// hf = true: the tokenizer of HF's CLIPTokenizer as the reference's eval path uses it (utils_attacks.py:67-71) - no
// html.unescape, HF's special-token spellings; control characters other than \t \n \r (which its BasicTokenizer drops
// instead of treating as white space) are outside the domain and flagged.
K1_HD int k1_prepare(const K1Tables& T, const uint8_t* src, int len, bool do_edit, int z, int c, K1Scratch& S, bool hf = false) {
  int flags = 0;
  int n = k1_decode_utf8(src, len, S.buf_b, &flags);       // code points, one byte each
  if (do_edit && (z < 0 || z > 2 * n)) { do_edit = false; flags |= K1_FLAG_TOO_LONG; }
  if (do_edit) n = k1_apply_edit(S.buf_b, n, z, c, S.buf_a);
  else for (int i = 0; i < n; ++i) S.buf_a[i] = S.buf_b[i];
  bool amp = false, lt = false;
  for (int i = 0; i < n; ++i) {
    const k1_char ch = S.buf_a[i];
    if (ch >= 0x80) {
      if (hf) flags |= K1_FLAG_NON_ASCII;                   // HF mode stays ASCII: BasicTokenizer drops U+00AD and other controls
      else if (ch <= 0x9F) flags |= K1_FLAG_NON_ASCII;      // C1 control: ftfy rewrites it as Windows-1252
      else if (ch >= 0xC2 && ch <= 0xF4 && i + 1 < n && k1_continuation_like(S.buf_a[i + 1]))
        flags |= K1_FLAG_NON_ASCII;                         // looks like UTF-8 read as Latin-1 / Windows-1252: ftfy re-decodes it
    }
    if (hf && (ch == 0x7f || (ch < 0x20 && ch != 9 && ch != 10 && ch != 13))) flags |= K1_FLAG_NON_ASCII;
    amp |= (ch == '&');
    lt |= (ch == '<');
  }
  if (hf) amp = false;
  const k1_char* cur = S.buf_a;
  if (amp) {                                               // html.unescape(html.unescape(text))
    n = k1_unescape(T, S.buf_a, n, S.buf_b, &flags, !lt);
    n = k1_unescape(T, S.buf_b, n, S.buf_a, &flags);
    if (!lt) {                                             // ftfy.fix_text unescapes one level first when the text has no '<':
      int f3 = 0;                                          // the reference then sees three levels; exact here iff the third is a no-op
      const int n3 = k1_unescape(T, S.buf_a, n, S.buf_b, &f3);
      bool same = n3 == n;
      for (int i = 0; same && i < n; ++i) same = S.buf_a[i] == S.buf_b[i];
      if (!same) flags |= K1_FLAG_ENTITY_DOMAIN;
    }
  }
  n = k1_clean(T, cur, n, S.buf_b, &flags);
  S.text_len = n;
  S.n_pieces = k1_split(T, S.buf_b, n, S.piece_start, S.piece_len, hf);
  return flags;
}

// ---- BPE of one piece (tokenizer.py:172-211 on integer ids; merged id = 512 + rank) ------------------------
// Runs on any lane; pieces own disjoint regions of sym / rk. piece_len[p] becomes the id count.
K1_HD void k1_encode_piece(const K1Tables& T, K1Scratch& S, int p) {
  const uint16_t pl = S.piece_len[p];
  const int start = S.piece_start[p];
  uint16_t* sym = S.sym + 2 * start;
  uint16_t* rk = S.rk + 2 * start;
  if (pl & 0x8000) {
    sym[0] = static_cast<uint16_t>((pl & 1) ? K1_EOT : K1_SOT);
    S.piece_len[p] = 1;
    return;
  }
  const k1_char* t = S.buf_b + start;
  int n = 0;
  for (int i = 0; i < pl; ++i) {                           // UTF-8 bytes of the code points (all < U+0800: 1 or 2 bytes) -> byte symbols
    const k1_char cp = t[i];
    if (cp < 0x80) sym[n++] = T.byte_id[cp];
    else { sym[n++] = T.byte_id[0xC0 | (cp >> 6)]; sym[n++] = T.byte_id[0x80 | (cp & 0x3F)]; }
  }
  sym[n - 1] = static_cast<uint16_t>(sym[n - 1] + 256);    // '</w>' on the last symbol (tokenizer.py:175)
  for (int i = 0; i + 1 < n; ++i) rk[i] = static_cast<uint16_t>(k1_merge_rank(T, sym[i], sym[i + 1]));
  while (n > 1) {
    uint32_t best = K1_NO_RANK;
    int bi = -1;
    for (int i = 0; i + 1 < n; ++i)
      if (rk[i] < best) { best = rk[i]; bi = i; }
    if (bi < 0) break;
    const uint16_t a = sym[bi], b = sym[bi + 1];
    const uint16_t merged = static_cast<uint16_t>(512u + best);
    int j = 0, i = 0;
    while (i < n) {                                        // every occurrence, left to right
      if (i + 1 < n && sym[i] == a && sym[i + 1] == b) {
        sym[j] = merged;
        if (j > 0) rk[j - 1] = K1_DIRTY;
        rk[j] = K1_DIRTY;
        ++j; i += 2;
      } else {
        sym[j] = sym[i];
        rk[j] = rk[i];                                     // pair (i, i+1) survives unless the next symbol merges
        ++j; ++i;
      }
    }
    n = j;
    for (int q = 0; q + 1 < n; ++q)
      if (rk[q] == K1_DIRTY) rk[q] = static_cast<uint16_t>(k1_merge_rank(T, sym[q], sym[q + 1]));
  }
  S.piece_len[p] = static_cast<uint16_t>(n);
}

// ---- row assembly (tokenizer.py:256-263): [SOT] + ids + [EOT], truncated to 77 with EOT forced, 0 padded --
// returns argmax(ids) + 1 (first EOT; transformer.py:661)
K1_HD int k1_emit_row(const K1Scratch& S, int32_t* row) {
  int k = 0;
  row[k++] = K1_SOT;
  for (int p = 0; p < S.n_pieces && k < K1_CTX - 1; ++p) {
    const uint16_t* sym = S.sym + 2 * S.piece_start[p];
    const int cnt = S.piece_len[p];
    for (int q = 0; q < cnt && k < K1_CTX - 1; ++q) row[k++] = sym[q];
  }
  row[k++] = K1_EOT;
  int first_eot = k - 1;
  for (int q = 1; q < k - 1; ++q)
    if (row[q] == K1_EOT) { first_eot = q; break; }
  for (; k < K1_CTX; ++k) row[k] = 0;
  return first_eot + 1;
}

}  // namespace leaf
