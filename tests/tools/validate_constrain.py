#!/usr/bin/env python
"""For a machine WITH NLTK and its data (nltk.download('words'), nltk.download('punkt')): measures how often the
restatement of the reference's `--constrain` filter (oracle/nltk_restate.py, which the CUDA kernel is pinned to) differs
from NLTK itself - the parity this repo could not pin offline. No GPU needed.

    python tests/tools/validate_constrain.py [captions.txt]        # one caption per line; default: built-in samples

It also replays tests/golden/nltk_published_vectors.json (NLTK's own docstring / doctest / unit-test examples, transcribed by
hand because NLTK is not installable in the build image) through the installed NLTK and reports every transcription error,
and writes the whole outcome to profiles/constrain_vs_nltk_report.txt so that it can be committed.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nltk_restate as N  # noqa: E402

try:
    import nltk
    from nltk.corpus import words
    from nltk.tokenize import word_tokenize
    W = frozenset(words.words())
    word_tokenize("a test.")
except Exception as ex:  # pragma: no cover
    sys.exit(f"NLTK with the 'words' and 'punkt' data is required for this check: {ex}")

try:
    abbrev = frozenset(nltk.data.load("tokenizers/punkt/english.pickle")._params.abbrev_types)
except Exception:
    abbrev = frozenset()

caps = [l.rstrip("\n") for l in open(sys.argv[1])] if len(sys.argv) > 1 else [
    "A big burly grizzly bear is show with grass in the background.", "a photo of a cat's toy, isn't it? Yes. The dog can't stop.",
    'He said "hello there." Then he left -- cannot go', "Mr. Smith's dog & cat: a,b at 3 p.m. in St. Louis"]
rng = random.Random(0)
V = [-1] + [ord(c) for c in "abcdefghijklmnopqrstuvwxyz ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~"]
from oracle.leaf_oracle import edit_sentence  # noqa: E402

report = []
pub = json.load(open(os.path.join(ROOT, "tests", "golden", "nltk_published_vectors.json")))
from nltk.tokenize import NLTKWordTokenizer, sent_tokenize  # noqa: E402
wrong = 0
for v in pub["vectors"]:
    got = {"treebank": lambda t: NLTKWordTokenizer().tokenize(t), "word_tokenize": word_tokenize, "sent_tokenize": sent_tokenize}[v["kind"]](v["text"])
    if got != v["expected"]:
        wrong += 1
        report.append(f"PUBLISHED VECTOR MISMATCH ({v['source']}): {v['text']!r}\n   nltk {nltk.__version__}: {got}\n   fixture: {v['expected']}")
report.append(f"published vectors: {len(pub['vectors']) - wrong}/{len(pub['vectors'])} reproduced by nltk {nltk.__version__}")

tot = tok_bad = cnt_bad = 0
for S in caps:
    for _ in range(200):
        s = edit_sentence(S, rng.randint(0, 2 * len(S)), rng.choice(V)).lower()
        a, b = word_tokenize(s), N.word_tokenize(s, abbrev)
        tot += 1
        tok_bad += a != b
        cnt_bad += len(W.intersection(a)) != len(W.intersection(b))
        if a != b and tok_bad <= 50:
            report.append(f"DIFF {s!r}\n   nltk: {a}\n   restatement: {b}")
report.append(f"{tot} candidates: token lists differ on {tok_bad} ({tok_bad / tot:.2%}), dictionary-word counts on {cnt_bad} ({cnt_bad / tot:.2%})")
print("\n".join(report))
out = os.path.join(ROOT, "profiles", "constrain_vs_nltk_report.txt")
open(out, "w").write("\n".join(report) + "\n")
print("written to", out)
