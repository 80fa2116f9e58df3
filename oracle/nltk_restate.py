"""CPU ORACLE for the reference's `--constrain` filter - TEST INFRASTRUCTURE ONLY.

    valid_sentence_batched(original, attacked)            /root/reference/utils_attacks.py:110-143
        W  = set(nltk.corpus.words.words())
        lo = len(W & set(word_tokenize(o.lower())))       per original sentence
        la = len(W & set(word_tokenize(a.lower())))       per candidate
        valid = la < lo

PARITY: PINNED TO NLTK'S PUBLISHED VECTORS ONLY. NLTK (unpinned in /root/reference/requirements.txt) and its `words` /
`punkt` data are not installed in this image and cannot be fetched, so this module was never run against a live NLTK.
What pins it: the 35 input/output pairs NLTK publishes in its own docstrings, doctests and unit tests
(tests/golden/nltk_published_vectors.json; 34 reproduced, the one known divergence is a Capitalised-text case outside
the reference's lower-cased domain - tests/test_constrain_cpu.py). The functions below RESTATE the published algorithm of
`nltk.word_tokenize` (NLTK 3.8.x):

  * `NLTKWordTokenizer.tokenize`  (nltk/tokenize/destructive.py): the regular-expression pipeline is written out below
    substitution by substitution, in NLTK's order, with Python's `re` - the same engine NLTK runs them on.
  * `sent_tokenize` (Punkt, nltk/tokenize/punkt.py) needs the trained English parameters (abbreviations, collocations,
    sentence starters, orthographic contexts), which only exist inside NLTK's data package. It is APPROXIMATED by
    `sent_split`: Punkt's first-pass rule (a token that ends in '.', is not an ellipsis and is not a known abbreviation
    ends a sentence; '?' and '!' always do) with a caller-supplied abbreviation set, plus the outcome of its second pass
    for lower-cased text (initials and numbers followed by a lower-case token do not end a sentence). Captions are
    almost always one sentence, where this step is the identity.
    Known, unmeasurable-here divergences from Punkt with english.pickle: (1) its trained collocations (`(typ, next_typ)`
    pairs that cancel a break) are not available; (2) for an initial or a number followed by a lower-case token whose
    orthographic context in the training corpus is "sentence-initial lower-case only", Punkt keeps the break and this
    module does not; (3) `!`/`?` inside a word followed by letters (`wh!mr. x`) count as breaks in Punkt's context
    tokenization and are ignored here; (4) the non-word character class is the pre-3.6.6 one (with `?`); (5) the
    abbreviation set is whatever the caller passes (engine.load_words(..., abbrev=...)), not english.pickle's.

The CUDA kernel (leaf_b200/csrc/constrain_core.cuh) is pinned bit-exactly against THIS module; this module is what a
user with NLTK installed should check first (tests/tools/validate_constrain.py does that and reports the mismatch rate).
"""
from __future__ import annotations

import re

# ---- nltk/tokenize/destructive.py :: NLTKWordTokenizer ---------------------------------------------------------------
STARTING_QUOTES = [
    (re.compile("([«“‘„]|[`]+)", re.U), r" \1 "),
    (re.compile(r"^\""), r"``"),
    (re.compile(r"(``)"), r" \1 "),
    (re.compile(r"([ \(\[{<])(\"|\'{2})"), r"\1 `` "),
    (re.compile(r"(?i)(\')(?!re|ve|ll|m|t|s|d|n)(\w)\b", re.U), r"\1 \2"),
]
ENDING_QUOTES = [
    (re.compile("([»”’])", re.U), r" \1 "),
    (re.compile(r"''"), " '' "),
    (re.compile(r'"'), " '' "),
    (re.compile(r"([^' ])('[sS]|'[mM]|'[dD]|') "), r"\1 \2 "),
    (re.compile(r"([^' ])('ll|'LL|'re|'RE|'ve|'VE|n't|N'T) "), r"\1 \2 "),
]
PUNCTUATION = [
    (re.compile(r'([^\.])(\.)([\]\)}>"\'' "»”’ " r"]*)\s*$", re.U), r"\1 \2 \3 "),
    (re.compile(r"([:,])([^\d])"), r" \1 \2"),
    (re.compile(r"([:,])$"), r" \1 "),
    (re.compile(r"\.{2,}", re.U), r" \g<0> "),
    (re.compile(r"[;@#$%&]"), r" \g<0> "),
    (re.compile(r'([^\.])(\.)([\]\)}>"\']*)\s*$'), r"\1 \2\3 "),
    (re.compile(r"[?!]"), r" \g<0> "),
    (re.compile(r"([^'])' "), r"\1 ' "),
    (re.compile(r"[*]", re.U), r" \g<0> "),
]
PARENS_BRACKETS = (re.compile(r"[\]\[\(\)\{\}\<\>]"), r" \g<0> ")
DOUBLE_DASHES = (re.compile(r"--"), r" -- ")
# nltk/tokenize/destructive.py :: MacIntyreContractions
CONTRACTIONS2 = [re.compile(p) for p in (
    r"(?i)\b(can)(?#X)(not)\b", r"(?i)\b(d)(?#X)('ye)\b", r"(?i)\b(gim)(?#X)(me)\b", r"(?i)\b(gon)(?#X)(na)\b",
    r"(?i)\b(got)(?#X)(ta)\b", r"(?i)\b(lem)(?#X)(me)\b", r"(?i)\b(more)(?#X)('n)\b", r"(?i)\b(wan)(?#X)(na)(?=\s)")]
CONTRACTIONS3 = [re.compile(p) for p in (r"(?i) ('t)(?#X)(is)\b", r"(?i) ('t)(?#X)(was)\b")]


def treebank_tokenize(text: str) -> list:
    """NLTKWordTokenizer().tokenize(text) with convert_parentheses=False."""
    for rx, sub in STARTING_QUOTES:
        text = rx.sub(sub, text)
    for rx, sub in PUNCTUATION:
        text = rx.sub(sub, text)
    text = PARENS_BRACKETS[0].sub(PARENS_BRACKETS[1], text)
    text = DOUBLE_DASHES[0].sub(DOUBLE_DASHES[1], text)
    text = " " + text + " "
    for rx, sub in ENDING_QUOTES:
        text = rx.sub(sub, text)
    for rx in CONTRACTIONS2:
        text = rx.sub(r" \1 \2 ", text)
    for rx in CONTRACTIONS3:
        text = rx.sub(r" \1 \2 ", text)
    return text.split()


# ---- sentence splitting: Punkt approximated (see the module docstring) ---------------------------------------------------
NONWORD = set("?!)\";}]*:@'({[")                      # punkt.py :: PunktLanguageVars._re_non_word_chars
_WS = set(" \t\n\r\x0b\x0c\x1c\x1d\x1e\x1f")           # str.isspace() over ASCII
_NUMBER = re.compile(r"^-?[\.,]?\d[\d,\.-]*$")         # punkt.py :: PunktToken._RE_NUMERIC without the final period
_CLOSERS = set("\"')]}")                              # punkt.py :: _re_boundary_realignment


def _is_end_context(text: str, i: int) -> bool:
    """punkt.py :: PunktLanguageVars._period_context_fmt at position i: a sentence-end character followed by a non-word
    character, or by whitespace and another token."""
    n = len(text)
    if text[i] not in ".?!" or i + 1 >= n:
        return False
    if text[i + 1] in NONWORD:
        return True
    j = i + 1
    while j < n and text[j] in _WS:
        j += 1
    return j > i + 1 and j < n


def sent_split(text: str, abbrev=frozenset()) -> list:
    """Sentence strings of `text` (already lower-cased by the caller, utils_attacks.py:132).

    Potential sentence ends that sit in the SAME whitespace-delimited word (`this!!! that`, `cat.! b`) are one decision, taken
    at the last of them (punkt.py :: PunktSentenceTokenizer._match_potential_end_contexts, NLTK >= 3.6.6: a match whose
    preceding-word slice overlaps the next match's is not yielded; the surviving context holds the whole word, so a
    sentence break found at any of its end characters counts)."""
    n = len(text)
    out, last = [], 0
    i = 0
    pending = False                                       # a break decided at an earlier end character of this word
    while i < n:
        c = text[i]
        if _is_end_context(text, i):
            nxt = text[i + 1]
            j = i + 1
            while j < n and text[j] in _WS:
                j += 1
            after_ws = j > i + 1 and j < n                # whitespace, then another token
            brk = True
            if c == ".":
                if (i > 0 and text[i - 1] == ".") or nxt == ".":
                    brk = False                       # ellipsis / multi-character punctuation
                else:
                    s = i
                    while s > 0 and text[s - 1] not in _WS and text[s - 1] not in NONWORD:
                        s -= 1
                    stem = text[s:i]
                    if stem:
                        if stem in abbrev or stem.split("-")[-1] in abbrev:
                            brk = False
                        elif len(stem) == 1 and stem.isalpha():
                            brk = False               # an initial followed by lower-case text
                        elif _NUMBER.match(stem):
                            brk = False               # a number / ordinal followed by lower-case text
            later = False                                 # another potential end further on in the same word?
            k = i + 1
            while k < n and text[k] not in _WS:
                if _is_end_context(text, k):
                    later = True
                    break
                k += 1
            if later:
                pending |= brk
                i += 1
                continue
            brk |= pending
            pending = False
            if brk:
                end = i + 1
                start = j if after_ws else i + 1
                # realign: closing quotes / brackets that open the next sentence belong to this one
                k = start
                while k < n and text[k] in _CLOSERS:
                    k += 1
                if k > start and (k == n or text[k] in _WS or text.startswith("--", k)):
                    end = k
                    while k < n and text[k] in _WS:
                        k += 1
                    start = k
                out.append(text[last:end])
                last = start
                i = max(i + 1, start)
                continue
        i += 1
    tail = text[last:].rstrip("".join(_WS))
    out.append(tail)
    return [s for s in out if s] or [""]


def word_tokenize(text: str, abbrev=frozenset()) -> list:
    """nltk.word_tokenize(text): Treebank tokens of every sentence."""
    return [tok for sent in sent_split(text, abbrev) for tok in treebank_tokenize(sent)]


def count_dictionary_words(text: str, words: frozenset, abbrev=frozenset()) -> int:
    """len(W.intersection(word_tokenize(text.lower())))  (utils_attacks.py:132,139)."""
    return len(words.intersection(word_tokenize(text.lower(), abbrev)))


def valid_sentence_batched(original, attacked, words: frozenset, abbrev=frozenset()):
    """utils_attacks.py:110-143 with the word list passed in (it is NLTK data): valid[b][j] = count(attacked[b][j]) <
    count(original[b])."""
    if isinstance(attacked, str):
        attacked = [[attacked]]
    if isinstance(attacked[0], str):
        attacked = [attacked]
    if isinstance(original, str):
        original = [original]
    lo = [count_dictionary_words(o, words, abbrev) for o in original]
    return [[count_dictionary_words(a, words, abbrev) < l for a in AS] for l, AS in zip(lo, attacked)]
