"""The on-device `--constrain` filter's scalar core (leaf_b200/csrc/constrain_core.cuh), compiled for the CPU by the test
harness, against the oracle's `re`-based restatement of nltk.word_tokenize (oracle/nltk_restate.py). Bit-exact: every
dictionary-word count must agree. NLTK is not installable here; what pins BOTH to NLTK are the input/output pairs NLTK itself
publishes in its docstrings, doctests and unit tests (tests/golden/nltk_published_vectors.json, transcribed by
tests/golden/make_nltk_published_vectors.py - none of those expected outputs came from this repo's code)."""
import json
import os
import random
import string

import numpy as np

from leaf_b200 import synth
from oracle import leaf_oracle as O
from oracle import nltk_restate as N
from tests import k1_harness as H

ABBREV = ["mr", "mrs", "dr", "st", "e.g", "i.e", "vs", "u.s", "a.m", "p.m", "inc", "no"]


def _word_list(seed=0):
    """A stand-in for nltk.corpus.words.words(): the pseudo-vocabulary of the synthetic captions plus common English words,
    contraction pieces and capitalised entries (which lower-cased tokens must never match)."""
    rng = random.Random(seed)
    vocab = synth._pseudo_vocab(seed)
    words = set(rng.sample(vocab, len(vocab) // 2))
    words |= {"a", "an", "the", "cat", "dog", "photo", "of", "can", "not", "do", "is", "it", "was", "more", "go", "stop", "he",
              "said", "hello", "there", "then", "left", "well", "known", "and", "or", "rock", "roll", "gim", "me", "gon", "na",
              "wan", "got", "ta", "lem", "d", "ye", "t", "s", "i", "you", "we", "they", "ll", "re", "ve", "Aaron", "A", "The"}
    return sorted(words)


def _texts(seed, count):
    rng = random.Random(seed)
    words = _word_list()
    frags = ["'s", "'t", "'re", "'ve", "'m", "'ll", "'d", "n't", "cannot", "gimme", "gonna", "gotta", "lemme", "wanna", "d'ye",
             "more'n", "'tis", "'twas", "...", "..", "--", "---", "``", "''", '"', "`", " . ", ". ", ".", "?", "!", "?!", ",", ":",
             ", ", ": ", ";", "(", ")", "[", "]", "{", "}", "<", ">", "*", "&", "@", "#", "$", "%", "mr. ", "e.g. ", "3.5", "3,000",
             "1. ", "a. ", "st. ", "-", "/", "_", "'", " '", "' ", "'a'", "'n ", "o'clock", "\t", "  "]
    out = []
    for _ in range(count):
        parts = []
        for _ in range(rng.randint(1, 12)):
            r = rng.random()
            if r < 0.55:
                parts.append(rng.choice(words).lower())
            elif r < 0.85:
                parts.append(rng.choice(frags))
            else:
                parts.append("".join(rng.choice(string.ascii_lowercase + string.digits + string.punctuation + "  ")
                                     for _ in range(rng.randint(1, 6))))
            if rng.random() < 0.7:
                parts.append(" ")
        out.append("".join(parts)[:300])
    return out


def test_counts_match_the_regex_restatement():
    words, abbrev = _word_list(), ABBREV
    H.load_words(words, abbrev)
    W, A = frozenset(words), frozenset(abbrev)
    texts = _texts(1, 4000) + ["", " ", ".", "a.", "a. b", '"a"', 'he said "hello there." then left', "don't stop. it's mr. smith's dog! really? yes.",
                               "cannot gimme gonna wanna go 'tis fine", "the cats' toys (red) & blue -- 3,000 items: a,b",
                               "well-known and/or 'a' rock 'n roll o'clock", "e.g. a cat. a dog.", "wait... what", "a 3. b", "x,", "x:",
                               "a 'twas b", "wanna", "wanna ", "i can't. you won't!", "a.)  b", 'a." b', "a.' --b", "end.)"]
    # through the candidate interface: every text is a "sentence" of its own with one no-op candidate
    pos = np.zeros((len(texts), 1), dtype=np.int32)
    chr_ = np.full((len(texts), 1), -1, dtype=np.int32)              # slot 0, delete marker = no-op
    counts, flags = H.constrain_counts(texts, 1, pos, chr_)
    assert flags == 0
    B = len(texts)
    for i, t in enumerate(texts):
        want = N.count_dictionary_words(t, W, A)
        assert counts[B + i] == want, (t, int(counts[B + i]), want, N.word_tokenize(t.lower(), A))
        assert counts[i] == want


def test_attack_shaped_masks_match_the_oracle():
    """Candidates as the attack makes them: one edit of V at a drawn position, both phases' forms."""
    words = _word_list()
    H.load_words(words, ABBREV)
    W, A = frozenset(words), frozenset(ABBREV)
    caps = synth.make_captions(12, seed=3) + ["a photo of a cat's toy, isn't it? yes. the dog (brown) can't stop.",
                                               'he said "hello there." then left -- cannot go', "mr. smith's dog & cat: a,b"]
    V = synth.V_DEFAULT
    rs = np.random.RandomState(0)
    n = 40
    pos = np.stack([rs.randint(0, 2 * len(S) + 1, size=n) for S in caps]).astype(np.int32)
    chr_ = np.asarray(V, dtype=np.int32)[rs.randint(0, len(V), size=(len(caps), n))]
    counts, flags = H.constrain_counts(caps, n, pos, chr_)
    assert flags == 0
    SS = [[O.edit_sentence(S, int(z), int(c)) for z, c in zip(pos[b], chr_[b])] for b, S in enumerate(caps)]
    want = N.valid_sentence_batched(caps, SS, W, A)
    B = len(caps)
    got = [[bool(counts[b * n + j] < counts[B * n + b]) for j in range(n)] for b in range(B)]
    assert got == want
    # with sel: every candidate of sample b sits at pos[b, sel[b]]
    sel = rs.randint(0, n, size=B).astype(np.int32)
    counts2, _ = H.constrain_counts(caps, n, pos, chr_, sel)
    SS2 = [[O.edit_sentence(S, int(pos[b, sel[b]]), int(c)) for c in chr_[b]] for b, S in enumerate(caps)]
    want2 = N.valid_sentence_batched(caps, SS2, W, A)
    assert [[bool(counts2[b * n + j] < counts2[B * n + b]) for j in range(n)] for b in range(B)] == want2
    assert 0.02 < np.mean(want) < 0.98          # the mask is not degenerate on this word list


# ---- NLTK's own published vectors --------------------------------------------------------------------------------------------
PUBLISHED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nltk_published_vectors.json")
# Vectors the restatement is KNOWN not to reproduce, with the reason (DESIGN.md section 2 lists them as well):
KNOWN_DIVERGENCES = {
    "1. This is R .\n2. This is A .\n3. That's all":
        "Punkt's orthographic heuristic breaks after a number when the next token is Capitalised; the reference lower-cases "
        "every sentence before tokenizing (utils_attacks.py:132,139), where the same rule keeps '1. this' together",
}


def published_vectors():
    return json.load(open(PUBLISHED))["vectors"]


def test_restatement_reproduces_nltk_published_vectors():
    """oracle/nltk_restate.py against outputs NLTK publishes for TreebankWordTokenizer / NLTKWordTokenizer.tokenize,
    word_tokenize and sent_tokenize. Sentence-level vectors are compared after lower-casing both sides (the reference's
    domain; the restatement's Punkt rules are the lower-case outcome of the orthographic heuristics)."""
    seen = {"treebank": 0, "word_tokenize": 0, "sent_tokenize": 0}
    diverged = []
    for v in published_vectors():
        ab = frozenset(v.get("abbrev", []))
        if v["kind"] == "treebank":
            got, want = N.treebank_tokenize(v["text"]), v["expected"]
        elif v["kind"] == "word_tokenize":
            got, want = N.word_tokenize(v["text"], ab), v["expected"]
            assert N.word_tokenize(v["text"].lower(), ab) == [t.lower() for t in want], v["text"]      # as the reference calls it
        else:
            got, want = N.sent_split(v["text"].lower(), ab), [t.lower() for t in v["expected"]]
        seen[v["kind"]] += 1
        if got != want:
            diverged.append(v["text"])
    assert seen["treebank"] >= 5 and seen["word_tokenize"] >= 15 and seen["sent_tokenize"] >= 14
    assert sorted(diverged) == sorted(KNOWN_DIVERGENCES), diverged


def published_count_cases():
    """(lower-cased text, abbreviations, tokens word_tokenize must produce, strings it must NOT produce) for every published
    vector that the filter's interface can express: ASCII text, expected tokens compared as a set."""
    out = []
    for v in published_vectors():
        if v["kind"] == "sent_tokenize" or not v["text"].isascii() or v["text"] in KNOWN_DIVERGENCES:
            continue
        if v["kind"] == "treebank" and any(t.endswith(".") and len(t) > 1 and set(t) != {"."} for t in v["expected"]):
            continue                          # Treebank alone keeps 'York.' inside a text; word_tokenize splits sentences first
        text = v["text"].lower()
        want = sorted({t.lower() for t in v["expected"]})
        glued = sorted({w for w in text.split() if w not in want})           # whitespace tokens that must have been split
        out.append((text, v.get("abbrev", []), want, glued))
    return out


def test_filter_core_counts_nltk_published_vectors():
    """The CPU-compiled CUDA core on NLTK's published word_tokenize vectors: with W = the published tokens (plus the glued
    whitespace forms as decoys that must never be counted... they are NOT in W, so a tokenizer that failed to split them
    would lose the words they contain) the count is exactly the number of distinct published tokens; with W = the glued
    forms only it is 0."""
    cases = published_count_cases()
    assert len(cases) >= 15
    for text, abbrev, want, glued in cases:
        pos, chr_ = np.zeros((1, 1), dtype=np.int32), np.full((1, 1), -1, dtype=np.int32)
        H.load_words(want, abbrev)
        counts, flags = H.constrain_counts([text], 1, pos, chr_)
        assert flags == 0 and counts[1] == len(want) and counts[0] == len(want), (text, int(counts[1]), want)
        if glued:
            H.load_words(glued, abbrev)
            counts, _ = H.constrain_counts([text], 1, pos, chr_)
            assert counts[1] == 0, (text, glued)
        for t in want:                                                            # and token by token
            H.load_words([t], abbrev)
            counts, _ = H.constrain_counts([text], 1, pos, chr_)
            assert counts[1] == 1, (text, t)
