"""The K1 integer core (leaf_b200/csrc/k1_core.cuh, compiled for the CPU by tests/k1_harness.py) against the
golden vectors generated from the reference and against the oracle. Bit-exact or fail."""
import json
import os
import random

import numpy as np

from leaf_b200 import synth
from oracle import leaf_oracle as O
from tests import k1_harness as H

V = synth.V_DEFAULT


MAX_CP = 0x24F


def _in_domain(text):
    """True when both html.unescape passes stay inside U+0000..U+024F (K1's domain) and str.lower() keeps every character of
    the result there (24 capitals of Latin Extended-B, and U+0130, do not)."""
    import html
    t1 = html.unescape(text)
    t2 = html.unescape(t1)
    return all(ord(c) <= MAX_CP for c in t1 + t2) and all(len(c.lower()) == 1 and ord(c.lower()) <= MAX_CP for c in t2)


def _row(ids):
    toks = [O.SOT] + list(ids) + [O.EOT]
    if len(toks) > 77:
        toks = toks[:77]
        toks[-1] = O.EOT
    return toks + [0] * (77 - len(toks))


def test_golden_tokenizer_strings(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "tokenizer_golden.json")))
    texts, want, unsup = [], [], []
    for text, ids in g["encode"]:
        if len(text) > 1000 or any(ord(c) > MAX_CP for c in text):
            continue
        (texts if _in_domain(text) else unsup).append(text)
        if texts and texts[-1] is text:
            want.append(_row(ids))
    assert len(texts) > 2500
    tok, ln, flags = H.expand_tokenize(texts)
    want = np.array(want, dtype=np.int32)
    bad = np.nonzero((tok != want).any(axis=1))[0]
    assert len(bad) == 0, [(texts[i], tok[i][:12].tolist(), want[i][:12].tolist()) for i in bad[:5]]
    assert (ln == want.argmax(axis=1) + 1).all()
    # entity expansions that leave U+0000..U+024F must be flagged, never silently mis-tokenized
    for t in unsup:
        _, _, fl = H.expand_tokenize([t])
        assert fl & 3, t
    rows_in = [r for r in g["rows_in"]]
    tok, ln, _ = H.expand_tokenize(rows_in)
    assert tok.tolist() == g["rows"]


def test_edit_rule_golden(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "edit_golden.json")))
    otok = O.OracleTokenizer()
    by_s = {}
    for S, z, c, out in g["cases"]:
        by_s.setdefault(S, []).append((z, c, out))
    for S, cases in by_s.items():
        n = len(cases)
        tok, ln, flags = H.expand_tokenize([S], n=n, pos=[c[0] for c in cases], chr_=[c[1] for c in cases])
        want = otok([c[2] for c in cases]).numpy()
        assert (tok == want).all(), S


def test_attack_shaped_candidates_vs_oracle():
    rng = random.Random(5)
    otok = O.OracleTokenizer()
    for kind, B in (("typical", 24), ("dense-77", 6), ("short", 12)):
        caps = synth.make_captions(B, seed=9, kind=kind)
        # salt a few captions with characters that make entities / contractions / specials reachable
        caps[0] = caps[0].replace(" ", " amp;", 1)
        caps[1] = "it" + caps[1][:20] + " dogs'street lt;b #65; not;in"
        caps[2] = "x start_of_text> y <end_of_text z " + caps[2]
        n = 64
        pos = np.array([[rng.randint(0, 2 * len(S)) for _ in range(n)] for S in caps], dtype=np.int32)
        chr_ = np.array([[V[rng.randrange(len(V))] for _ in range(n)] for _ in caps], dtype=np.int32)
        chr_[:, :8] = [ord(c) for c in "&<'_; #x"]
        valid = np.ones((B, n), dtype=np.uint8)
        valid[:, 5::7] = 0
        # phase-1 form
        tok, ln, flags = H.expand_tokenize(caps, n=n, pos=pos, chr_=chr_, valid=valid)
        strings = [O.edit_sentence(S, int(pos[b, j]), int(chr_[b, j])) if valid[b, j] else S
                   for b, S in enumerate(caps) for j in range(n)]
        want = otok(strings).numpy()
        ok = np.array([_in_domain(s) for s in strings])
        assert ok.mean() > 0.95
        assert (tok[ok] == want[ok]).all()
        assert (ln[ok] == want[ok].argmax(axis=1) + 1).all()
        # phase-2 form: one position per sample selected on the device side
        sel = np.array([rng.randrange(n) for _ in caps], dtype=np.int32)
        tok, ln, flags = H.expand_tokenize(caps, n=n, pos=pos, chr_=chr_, sel=sel)
        strings = [O.edit_sentence(S, int(pos[b, sel[b]]), int(chr_[b, j])) for b, S in enumerate(caps) for j in range(n)]
        want = otok(strings).numpy()
        ok = np.array([_in_domain(s) for s in strings])
        assert (tok[ok] == want[ok]).all()


def test_random_fuzz_vs_oracle():
    rng = random.Random(17)
    otok = O.OracleTokenizer()
    alpha = "abcdefghijklmnopqrstuvwxyz" * 2 + "ABCXYZ0123456789" + "    ''&&;;<>_#xX.,!?-" + '"$%()*+/:=@[\\]^`{|}~'
    frags = ["&lt", "&gt;", "&amp", "&not", "&copy", "&nbsp", "&#", "&#x", "&#65", "&#x41;", "'s", "'re", "'ll",
             "<start_of_text>", "<end_of_text>", "&amp;lt;", "&ampamp;", "&shy", "&micro", "&frac12", "&#32;", "&#9"]
    texts = []
    for _ in range(4000):
        s = "".join(rng.choice(alpha) for _ in range(rng.randint(0, 60)))
        for _ in range(rng.randint(0, 3)):
            p = rng.randint(0, len(s))
            s = s[:p] + rng.choice(frags) + s[p:]
        texts.append(s)
    texts += ["a" * 1000, "a1" * 500, ("&lt" * 300)[:1000], " " * 1000, "&" * 1000, "'" * 999 + "s"]
    tok, ln, flags = H.expand_tokenize(texts)
    want = otok(texts).numpy()
    ok = np.array([_in_domain(s) for s in texts])
    assert ok.mean() > 0.9
    bad = np.nonzero((tok != want).any(axis=1) & ok)[0]
    assert len(bad) == 0, [(texts[i], tok[i][:10].tolist(), want[i][:10].tolist()) for i in bad[:5]]


def test_hf_mode_matches_transformers(golden_dir):
    """The K1 core in HF mode (leaf_set_tokenizer_mode 1) against transformers' CLIPTokenizer."""
    import json
    g = json.load(open(os.path.join(golden_dir, "hf_tokenizer_golden.json")))
    texts = [s for s, _ in g["encode"] if len(s) <= 1000]
    tok, ln, flags = H.expand_tokenize(texts, hf=True)
    assert flags == 0
    for i, (s, ids) in enumerate((s, ids) for s, ids in g["encode"] if len(s) <= 1000):
        want = ids if len(ids) <= 77 else ids[:76] + [49407]
        assert tok[i, :len(want)].tolist() == want, s
        assert not tok[i, len(want):].any()
        assert ln[i] == want.index(49407) + 1                    # pooled position: the FIRST EOS (argmax)
    H.expand_tokenize(["x"])                                  # back to the default mode for the other tests


_CONT_LIKE = set(range(0x80, 0xC0)) | {0x152, 0x153, 0x160, 0x161, 0x178, 0x17D, 0x17E, 0x192}


def _ftfy_risk(text):
    """The inputs K1 flags because the real ftfy.fix_text would rewrite them (csrc/k1_core.cuh: domain): C1 controls, and a
    UTF-8-lead-like character followed by what a continuation byte looks like after a wrong Latin-1 / Windows-1252 decode."""
    for i, ch in enumerate(text):
        o = ord(ch)
        if 0x80 <= o <= 0x9F:
            return True
        if 0xC2 <= o <= 0xF4 and i + 1 < len(text) and ord(text[i + 1]) in _CONT_LIKE:
            return True
    return False


def test_latin1_utf8_captions_golden(golden_dir):
    """UTF-8 captions with code points up to U+024F (accented Western and Central European text): the reference's SimpleTokenizer on the
    captions and on attack-shaped edits whose positions count CODE POINTS (tests/golden/tokenizer_latin1_golden.json, generated
    from the reference by oracle/make_golden.py latin1); the domain flags for everything beyond."""
    g = json.load(open(os.path.join(golden_dir, "tokenizer_latin1_golden.json")))
    texts = [t for t, _ in g["encode"]]
    assert sum(not t.isascii() for t in texts) >= 28 and sum(max(map(ord, t)) > 0xFF for t in texts if t) >= 9
    tok, ln, flags = H.expand_tokenize(texts)
    want = np.array([_row(ids) for _, ids in g["encode"]], dtype=np.int32)
    assert (tok == want).all(), [t for t, a, b in zip(texts, tok, want) if (a != b).any()]
    assert (ln == want.argmax(axis=1) + 1).all()
    assert tok.tolist() == g["rows"] and flags == 0
    by_s = {}
    for S, z, c, out, ids in g["edits"]:
        by_s.setdefault(S, []).append((z, c, out, ids))
    n_risky = 0
    for S, cases in by_s.items():
        tok, ln, flags = H.expand_tokenize([S], n=len(cases), pos=[c[0] for c in cases], chr_=[c[1] for c in cases])
        for j, (z, c, out, ids) in enumerate(cases):
            assert out == O.edit_sentence(S, z, c)
            assert tok[j].tolist() == _row(ids), (S, z, c, out)
            _, _, fl = H.expand_tokenize([out])
            n_risky += _ftfy_risk(out)
            assert bool(fl & 2) == _ftfy_risk(out), (out, fl)
    # outside the domain: loud, never a silently different row
    for bad, bit in (("na\u00efve \u0250", 2), ("caf\u00e9 \u2019s", 2), ("emoji \U0001F600", 2), ("a\u0085b", 2), ("Ã©t\u00e9", 2),
                     ("\u0130stanbul", 2), ("\u023a", 2), ("Å¡koda", 2), ("Ã\u0153", 2), ("&#304;", 2), ("&#592;", 1),
                     ("x &amp;amp;lt; y", 1), ("x &EACUTE; y", 1), ("ok &amp;lt; <b>", 0), ("caf\u00e9 &eacute;", 0), ("\u0141\u00f3d\u017a", 0)):
        _, _, fl = H.expand_tokenize([bad])
        assert (fl & 3) == bit, (bad, fl)
    # invalid UTF-8 bytes are flagged as well
    data = np.frombuffer(b"ab\xff\xc3" + b"\0", dtype=np.uint8).copy()
    off = np.array([0, 4], dtype=np.int32)
    L = H.lib()
    import ctypes
    t, l = np.zeros((1, 77), dtype=np.int32), np.zeros(1, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert L.k1h_expand_tokenize(p(data), p(off), 1, 0, None, None, None, None, p(t), p(l)) & 2


def test_long_captions_up_to_3560_bytes():
    """Captions beyond the default variant's 1000 bytes (the reference's tokenizer takes any length, tokenizer.py:226-265): the
    long-text buffers of the kernel's second variant, against the oracle tokenizer, with edits anywhere in the text - most of
    them behind the 77-token window, where the row must equal the caption's; longer than 3560 bytes is flagged."""
    rng = random.Random(3)
    otok = O.OracleTokenizer()
    caps = [" ".join(synth.make_captions(40, seed=s, kind="typical"))[:L] for s, L in ((1, 1500), (2, 2600), (3, 3560))]
    caps += ["x" * 3560, ("ab " * 1400)[:3500], "caf\u00e9 " * 500, "a" + " " * 3000 + "b &amp; c"]
    assert max(len(c.encode("utf-8")) for c in caps) == 3560
    n = 24
    pos = np.array([[rng.randint(0, 2 * len(S)) for _ in range(n)] for S in caps], dtype=np.int32)
    pos[:, :4] = [[0, 1, 40, 90]] * len(caps)                                # some inside the window
    chr_ = np.array([[V[rng.randrange(len(V))] for _ in range(n)] for _ in caps], dtype=np.int32)
    tok, ln, flags = H.expand_tokenize(caps, n=n, pos=pos, chr_=chr_)
    assert flags == 0
    strings = [O.edit_sentence(S, int(pos[b, j]), int(chr_[b, j])) for b, S in enumerate(caps) for j in range(n)]
    want = otok(strings).numpy()
    assert (tok == want).all()
    base, _, _ = H.expand_tokenize(caps)
    assert (base == otok(caps).numpy()).all()
    _, _, fl = H.expand_tokenize(["y" * 3561])
    assert fl & 4
