#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY. Needs /root/reference (absent on the GPU box - the fixtures it
writes are committed so nothing at test time reads the reference tree).

    python oracle/make_golden.py

Imports, through oracle/ref_shims (ftfy = identity, nltk/torchmetrics stubs):
  * utils_attacks.generate_sentence / attack_text_leaf      (/root/reference/utils_attacks.py)
  * open_clip.tokenizer.SimpleTokenizer                     (/root/reference/src/open_clip/tokenizer.py)
  * open_clip.model.CLIP.encode_text                        (/root/reference/src/open_clip/model.py)
"""
import json
import os
import random
import string
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LEAF_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "ref_shims"), os.path.join(REF, "src"), REF, ROOT]

import open_clip  # noqa: E402
import utils_attacks  # noqa: E402
from open_clip.model import CLIP, CLIPTextCfg, CLIPVisionCfg  # noqa: E402

from leaf_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
V = synth.V_DEFAULT


def gen_edit():
    rng = random.Random(7)
    strings = ["", "a", "ab", "a b", "hello world", "a_b", "  x ", "Cat's", "&lt", "zz top!"]
    for _ in range(10):
        n = rng.randint(3, 14)
        strings.append("".join(rng.choice(string.ascii_letters + "  _&;'") for _ in range(n)))
    chars = [0, 27, 1, 2, 28, 53, 60, 70, 95] + [V.index(ord(c)) for c in "_&;' "]
    cases = []
    for S in strings:
        for z in range(2 * len(S) + 1):
            us = set(chars)
            if z % 2 == 1 and ord(S[z // 2]) in V:
                us.add(V.index(ord(S[z // 2])))       # "same character" => delete
            for u in sorted(us):
                out = utils_attacks.generate_sentence(S, z, u, V, 1, alternative=-1)
                cases.append([S, z, V[u], out])
    # the phase-1 form: generate_all_sentences(S, [ord(' ')], subset_z, alternative=-1)
    probe = []
    for S in strings[:12]:
        zs = list(range(2 * len(S) + 1))
        probe.append([S, zs, utils_attacks.generate_all_sentences(S, [ord(" ")], subset_z=zs, alternative=-1)])
    json.dump({"V": V, "cases": cases, "probe": probe}, open(os.path.join(OUT, "edit_golden.json"), "w"))
    print("edit cases", len(cases))


APPENDIX_B = [
    "a photo of a cat", "it's", "dogs'street", "a!!'s", "3d 42", "a_b", "x <start_of_text> y", "&amp;",
    "a &#65; b", "x&ltb", "cats &not here", "AT&T center &cent", "  Hello   WORLD  ", "e.g. a cat.",
    "don't'll", "#$%&'()*+", "a\tb\nc", "a" * 90, "a" * 90 + " " + "zq " * 80,
    "&not &notit; &lt &amp;lt; &#x41 &#65 &#0; &#128; &copy", "", " ", "&", "&;", "&#", "&#;", "&#x;", "&#xg",
    "<end_of_text>", "a<end_of_text>b", "!<start_of_text>", "<START_OF_TEXT>", "&amplt", "&ampamp;lt",
    "&nbsp", "a&nbspb", "&#9;x", "a&#32;b", "&#10", "&shy", "x&shyy", "&micro", "&MICRO", "&Aacute", "&aacute",
    "&AMP", "&QUOT", "&quot", "&THORN", "&eth", "&ETH", "&times", "&divide", "&frac12", "&sup2", "1&frac123",
    "&#255", "&#256", "&#xff;", "&#XFF", "&#x100", "&#65x", "&#0065", "&#1 dad", "&#128", "&#150;", "&#x9f",
    "&lambda;", "&Lambda;", "&hellip;", "&notin;", "&notin", "&lt;&gt;", "&lt&gt", "a&b", "a & b", "&&lt",
    "&#38;lt;", "&#38lt", "&amp;#65;", "&ordf", "&ordm", "&szlig", "&yuml", "&sup1&sup3", "&para&sect",
    "'S", "'RE", "A'LL b'D", "''s", "'s's", "x'", "'", "1'2", "a1b2", "12ab", "a.b", "...", "a - b",
    "don’t", "\x1c", "a\x1cb", "a\x85b", "a\xa0b", "\x0b\x0c",
]


def gen_tokenizer():
    tok = open_clip.get_tokenizer("ViT-L-14")
    rng = random.Random(11)
    texts = list(APPENDIX_B)
    alpha = string.ascii_lowercase * 3 + string.ascii_uppercase + string.digits + "   '&;<_#x" + string.punctuation
    frags = ["&lt", "&gt", "&amp", "&not", "&copy", "&reg", "&deg", "&cent", "&nbsp", "&#", "&#x", "'s", "'t",
             "'re", "'ve", "'m", "'ll", "'d", "<start_of_text>", "<end_of_text>", ";", "&quot", "&yen", "&uml",
             "&para", "&sect", "&shy", "&eth", "&times", "&micro", "&frac12", "&Aacute", "&ntilde", "&#65", "&#x41"]
    for _ in range(2600):
        n = rng.randint(1, 70)
        s = "".join(rng.choice(alpha) for _ in range(n))
        for _ in range(rng.randint(0, 3)):
            p = rng.randint(0, len(s))
            s = s[:p] + rng.choice(frags) + s[p:]
        texts.append(s)
    # attack-shaped strings: synthetic captions with one edit from V
    caps = synth.make_captions(24, seed=3, kind="typical") + synth.make_captions(4, seed=3, kind="dense-77")
    for S in caps:
        for _ in range(12):
            z = rng.randint(0, 2 * len(S))
            u = rng.randrange(len(V))
            texts.append(utils_attacks.generate_sentence(S, z, u, V, 1, alternative=-1))
    enc = [[t, tok.encode(t)] for t in texts]
    rows_in = ["a photo of a cat", "a" * 90 + " " + "zq " * 80, "", caps[-1], caps[0], "x <end_of_text> y"]
    rows = tok(rows_in).tolist()
    json.dump({"encode": enc, "rows_in": rows_in, "rows": rows},
              open(os.path.join(OUT, "tokenizer_golden.json"), "w"))
    print("tokenizer strings", len(enc))


LATIN1_CAPTIONS = [
    "un café crème à la française", "Ein schöner Tag in München mit Straße und Fluß", "El niño comió jalapeños en Ávila",
    "ÉCOLE SUPÉRIEURE DE LYON À PARIS", "Ærøskøbing er en by på Ærø", "naïve façade coöperate résumé", "señor Ñandú's piñata isn't here",
    "µm and ªº ordinals 1º 2ª", "½ cup ¼ tsp ¾ oz ¹²³ x²", "3×4÷2 ±1 °C §5 ¶ © ® «quoted» ¿qué? ¡hola!", "ÿ y-less þorn Þing ðat Ðe",
    "a\u00a0b non\u00a0breaking  spaces\u00a0", "Crème brûlée's déjà-vu: voilà!it's", "ß", "É", "éé 12½x ·middle· ¬not ¦bar ¨uml ¯macr ´acute ¸cedil",
    "soft\u00adhyphen in\u00adside", "£5 ¥6 ¢7 ¤8 and 9€-less", "Ò Ó Ô Õ Ö Ø Ù Ú Û Ü Ý à á â ã ä å æ ç è é ê ë ì í î ï ð ñ ò ó ô õ ö ø ù ú û ü ý þ ÿ",
    "À Á Â Ã Ä Å Æ Ç È É Ê Ë Ì Í Î Ï Ð Ñ", "façade&amp;café &eacute;t&eacute; &Eacute;T&Eacute; &#233; &#xE9;", "l'été d'août m'a plu 's 't 're",
    # Latin Extended-A / -B (U+0100..U+024F): Central European, Baltic, Romanian, Turkish (without U+0130), Vietnamese base letters
    "Zażółć gęślą jaźń w Łodzi", "Příliš žluťoučký kůň úpěl ďábelské ódy", "Árvíztűrő tükörfúrógép ŐŰ", "ȘTEFAN și ȚARA: mâță în București",
    "Ğğ Şş ı dotless and ǅ ǆ Ǆ digraphs Ǉǈǉ", "Ēēģīķļņū Šš Žž Ąą Ęę Ėė Įį Ųų", "\u017fhort long-s it'\u017f <\u017ftart_of_text> x", "Œuvre cœur Ÿ ÿ ƒ Ǝ ǝ Ȝ ȝ",
    "&Scaron;koda &zcaron;ena &#256;&#257; &#x17D; &OElig;&oelig; &Yuml; &fnof;", "ĀĂĄĆĈĊČĎĐĒĔĖĘĚĜĞĠĢĤĦĨĪĬĮĲĴĶĹĻĽĿŁŃŅŇŊŌŎŐŒŔŖŘŚŜŞŠŢŤŦŨŪŬŮŰŲŴŶŸŹŻŽ",
]


def gen_tokenizer_latin1():
    """Captions with code points in U+0080..U+024F (Latin-1 Supplement, Latin Extended-A / -B: what K1 accepts as UTF-8 input
    beyond ASCII): the reference's
    SimpleTokenizer on the captions and on attack-shaped edits of them (utils_attacks.generate_sentence, positions counted in
    code points). ftfy is the identity shim here, as for every other fixture; K1 flags the Latin-1 inputs on which the real
    ftfy would not be (C1 controls, mojibake-shaped pairs) and none of these captions is one of them, except the pairs an
    edit can create by deleting the character between a lead-like and a continuation-like one - those cases are kept in the
    fixture (their token ids are still what SimpleTokenizer produces under the shim) and marked."""
    tok = open_clip.get_tokenizer("ViT-L-14")
    rng = random.Random(23)
    caps = [c for c in LATIN1_CAPTIONS if c.isascii() or max(map(ord, c)) <= 0x24F]
    assert len(caps) == len(LATIN1_CAPTIONS) - 1                                   # the one with the euro sign is dropped
    enc = [[t, tok.encode(t)] for t in caps]
    edits = []
    for S in caps:
        for _ in range(16):
            z = rng.randint(0, 2 * len(S))
            u = rng.randrange(len(V))
            out = utils_attacks.generate_sentence(S, z, u, V, 1, alternative=-1)
            edits.append([S, z, V[u], out, tok.encode(out)])
    rows = tok(caps).tolist()
    json.dump({"encode": enc, "edits": edits, "rows_in": caps, "rows": rows},
              open(os.path.join(OUT, "tokenizer_latin1_golden.json"), "w"))
    print("latin-1 tokenizer strings", len(enc), "edits", len(edits))


def build_ref_clip(cfg: synth.TowerCfg, sd):
    model = CLIP(embed_dim=cfg.embed_dim,
                 vision_cfg=CLIPVisionCfg(layers=1, width=64, head_width=32, patch_size=16, image_size=32),
                 text_cfg=CLIPTextCfg(context_length=cfg.context_length, vocab_size=cfg.vocab_size,
                                      width=cfg.width, heads=cfg.heads, layers=cfg.layers),
                 quick_gelu=cfg.quick_gelu)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(m.startswith("visual.") or m in ("logit_scale",) for m in missing), missing
    return model.eval()


def gen_tower():
    tok = open_clip.get_tokenizer("ViT-L-14")
    caps = synth.make_captions(10, seed=5, kind="typical") + synth.make_captions(2, seed=5, kind="dense-77") \
        + ["", "a", "x <end_of_text> y z"]
    tokens = tok(caps)
    out = {"tokens": tokens.numpy()}
    for name, quick in (("tiny", False), ("tiny", True), ("small", False)):
        cfg = synth.TOWERS[name]
        cfg = synth.TowerCfg(cfg.name, cfg.width, cfg.layers, cfg.heads, cfg.embed_dim, quick_gelu=quick)
        sd = synth.random_tower_state_dict(cfg, seed=21, exact_numpy=True)
        model = build_ref_clip(cfg, sd)
        with torch.no_grad():
            f = model.encode_text(tokens, normalize=False)
            fn = model.encode_text(tokens, normalize=True)
        tag = f"{name}_{'quick' if quick else 'gelu'}"
        out[tag] = f.numpy()
        out[tag + "_norm"] = fn.numpy()
    np.savez_compressed(os.path.join(OUT, "tower_golden.npz"), **out)
    json.dump({"captions": caps}, open(os.path.join(OUT, "tower_golden.json"), "w"))
    print("tower fixtures", {k: v.shape for k, v in out.items()})


def gen_convert_ids():
    """The one known-answer input the reference's own tower checks use (conversion/convert_2.py:252-253 and
    convert_to_openclip.py:155-156 compare open_clip with HF on ids [49406, 1, ..., 76] at atol 1e-4): the row has no
    end-of-text token, so argmax(ids) pools position 0, the start-of-text token. Plus rows whose maximum sits first /
    last / in the middle of a full 77-token row."""
    ids = np.zeros((4, 77), dtype=np.int64)
    ids[0] = [49406] + list(range(1, 77))
    ids[1] = list(range(1, 77)) + [49407]
    ids[2] = [49406] + list(range(100, 137)) + [49407] + list(range(200, 238))
    ids[3] = [49406, 320, 49407] + [0] * 74
    tokens = torch.from_numpy(ids)
    out = {"tokens": ids}
    for name in ("tiny", "small"):
        cfg = synth.TOWERS[name]
        sd = synth.random_tower_state_dict(cfg, seed=23, exact_numpy=True)
        model = build_ref_clip(cfg, sd)
        with torch.no_grad():
            out[name] = model.encode_text(tokens, normalize=False).numpy()
    np.savez_compressed(os.path.join(OUT, "convert_ids_golden.npz"), **out)
    print("convert-ids fixtures", {k: v.shape for k, v in out.items()})


def gen_attack():
    tok = open_clip.get_tokenizer("ViT-L-14")
    cfg = synth.TOWERS["tiny"]
    sd = synth.random_tower_state_dict(cfg, seed=31, exact_numpy=True)
    sd_frozen = synth.perturbed_copy(sd, seed=32, std=1e-2, exact_numpy=True)
    model = build_ref_clip(cfg, sd)
    frozen = build_ref_clip(cfg, sd_frozen)
    cases = []
    arrays = {}
    for ci, (B, n, k, seed, kind, objective) in enumerate([
            (6, 10, 1, 0, "typical", "l2"), (6, 50, 1, 1, "typical", "l2"), (4, 120, 1, 2, "short", "l2"),
            (5, 20, 2, 3, "typical", "l2"), (3, 16, 3, 4, "short", "l2"), (2, 30, 1, 5, "dense-77", "l2"),
            (1, 24, 1, 6, "typical", "sim"), (1, 24, 1, 7, "typical", "dissim"), (4, 24, 1, 8, "typical", "negl2")]):
        caps = synth.make_captions(B, seed=100 + seed, kind=kind)
        with torch.no_grad():
            anchor = frozen.encode_text(tok(caps), normalize=objective in ("sim", "dissim"))
            np.random.seed(seed)
            feats, adv = utils_attacks.attack_text_leaf(model, tok, caps, anchor.clone(), "cpu", objective=objective,
                                                        n=n, k=k, V=V, constrain=False)
        cases.append(dict(B=B, n=n, k=k, seed=seed, kind=kind, objective=objective, captions=caps, adv=adv))
        arrays[f"anchor_{ci}"] = anchor.numpy()
        arrays[f"feats_{ci}"] = feats.numpy()
    json.dump({"tower": "tiny", "seed": 31, "frozen_seed": 32, "frozen_std": 1e-2, "cases": cases},
              open(os.path.join(OUT, "attack_golden.json"), "w"))
    np.savez_compressed(os.path.join(OUT, "attack_golden.npz"), **arrays)
    print("attack cases", len(cases))


def gen_attack_latin():
    """attack_text_leaf of the reference on ACCENTED captions (code points up to U+024F, positions drawn over code points):
    the whole attack on the K1 kernel's widened input domain, k = 1 and k = 2."""
    tok = open_clip.get_tokenizer("ViT-L-14")
    cfg = synth.TOWERS["tiny"]
    sd = synth.random_tower_state_dict(cfg, seed=31, exact_numpy=True)
    sd_frozen = synth.perturbed_copy(sd, seed=32, std=1e-2, exact_numpy=True)
    model = build_ref_clip(cfg, sd)
    frozen = build_ref_clip(cfg, sd_frozen)
    pool = [c for c in LATIN1_CAPTIONS if not c.isascii() and max(map(ord, c)) <= 0x24F and "&" not in c and "\u017f" not in c and len(c) > 12]
    cases, arrays = [], {}
    for ci, (lo, hi, n, k, seed) in enumerate([(0, 6, 24, 1, 0), (6, 12, 40, 1, 1), (12, 17, 16, 2, 2)]):
        caps = pool[lo:hi]
        with torch.no_grad():
            anchor = frozen.encode_text(tok(caps), normalize=False)
            np.random.seed(seed)
            feats, adv = utils_attacks.attack_text_leaf(model, tok, caps, anchor.clone(), "cpu", objective="l2", n=n, k=k, V=V,
                                                        constrain=False)
        cases.append(dict(B=len(caps), n=n, k=k, seed=seed, objective="l2", captions=caps, adv=adv))
        arrays[f"anchor_{ci}"] = anchor.numpy()
        arrays[f"feats_{ci}"] = feats.numpy()
    json.dump({"tower": "tiny", "seed": 31, "frozen_seed": 32, "frozen_std": 1e-2, "cases": cases},
              open(os.path.join(OUT, "attack_latin_golden.json"), "w"))
    np.savez_compressed(os.path.join(OUT, "attack_latin_golden.npz"), **arrays)
    print("accented attack cases", len(cases), [c["B"] for c in cases])


def gen_eval_attacks():
    """attack_text_charmer_inference / attack_text_bruteforce (utils_attacks.py:395-580), one sentence at a time."""
    tok = open_clip.get_tokenizer("ViT-L-14")
    cfg = synth.TOWERS["tiny"]
    sd = synth.random_tower_state_dict(cfg, seed=31, exact_numpy=True)
    sd_frozen = synth.perturbed_copy(sd, seed=32, std=1e-2, exact_numpy=True)
    sd2 = synth.random_tower_state_dict(cfg, seed=33, exact_numpy=True)
    model, frozen, model2 = build_ref_clip(cfg, sd), build_ref_clip(cfg, sd_frozen), build_ref_clip(cfg, sd2)
    charmer, brute, arrays = [], [], {}
    caps = synth.make_captions(3, seed=200, kind="typical") + synth.make_captions(2, seed=201, kind="short") + ["a cat's toy & a dog", "ab"]
    for ci, (si, n, k, objective, bs, two) in enumerate([
            (0, 10, 1, "l2", 2560, False), (1, 5, 2, "l2", 64, False), (2, 10, 1, "negl2", 2560, False),
            (3, 20, 3, "l2", 2560, False), (4, 3, 1, "sim", 2560, False), (5, 10, 1, "dissim", 100, False),
            (6, 10, 1, "l2", 2560, False), (0, 6, 1, "dissim", 2560, True), (5, 4, 2, "sim", 2560, True),
            (5, 4, 2, "dissim", 2560, True), (1, 8, 2, "negl2", 2560, True)]):
        S = caps[si]
        norm = objective in ("sim", "dissim")
        with torch.no_grad():
            anchor = frozen.encode_text(tok([S]), normalize=norm)
            anchor2 = model2.encode_text(tok(["another " + S]), normalize=norm) if two else None
            adv, dist = utils_attacks.attack_text_charmer_inference(
                model, tok, S, anchor.clone(), "cpu", objective=objective, n=n, k=k, V=V, constrain=False, batch_size=bs,
                model_2=model2 if two else None, model_2_anchor_features=anchor2.clone() if two else None)
        # torch.topk / argmax on exactly equal scores (no-op edits all score like the sentence itself) make the
        # reference's own result depend on the tie order of the torch build: flag those runs
        from oracle import leaf_oracle as O
        otok = O.OracleTokenizer()
        enc = lambda t, normalize, sd_=sd: O.encode_text(sd_, t, cfg.heads, normalize=normalize)
        enc2 = lambda t, normalize: O.encode_text(sd2, t, cfg.heads, normalize=normalize)
        trace = {}
        with torch.no_grad():
            oadv, _ = O.attack_text_charmer_oracle(enc, otok, S, anchor.clone(), objective=objective, n=n, k=k, batch_size=bs,
                                                   encode_2=enc2 if two else None,
                                                   anchor_2=anchor2.clone() if two else None, trace=trace)
        assert oadv == adv
        tie = False
        for r in trace["rounds"]:
            v = torch.sort(r["loss1"], descending=True).values
            kk = len(r["top"])
            tie |= kk < len(v) and bool(v[kk - 1] == v[kk])      # WHICH positions are kept (their order only matters for a tied maximum)
            # torch.argmax returns the FIRST maximum, so a tied maximum only depends on the torch build when the tied
            # candidates sit at different kept positions whose own order was a tie
            groups = set((torch.nonzero(r["loss2"] == r["loss2"].max()).flatten() // len(V)).tolist())
            tie |= len(groups) > 1 and bool((v[:kk - 1] == v[1:kk]).any())
        charmer.append(dict(sentence=S, n=n, k=k, objective=objective, batch_size=bs, two=two, adv=adv, dist=dist,
                            tie_dependent=tie))
        arrays[f"charmer_anchor_{ci}"] = anchor.numpy()
        if two:
            arrays[f"charmer_anchor2_{ci}"] = anchor2.numpy()
    for ci, (si, objective, bs) in enumerate([(3, "l2", 2560), (4, "dissim", 500), (6, "l2", 64), (1, "l2", 2560)]):
        S = caps[si]
        with torch.no_grad():
            anchor = frozen.encode_text(tok([S]), normalize=objective == "dissim")
            adv, dist = utils_attacks.attack_text_bruteforce(model, tok, S, anchor.clone(), "cpu", batch_size=bs,
                                                             objective=objective, V=V, constrain=False)
        brute.append(dict(sentence=S, objective=objective, batch_size=bs, adv=adv, dist=dist))
        arrays[f"brute_anchor_{ci}"] = anchor.numpy()
    json.dump({"tower": "tiny", "seed": 31, "seed2": 33, "charmer": charmer, "bruteforce": brute},
              open(os.path.join(OUT, "eval_attack_golden.json"), "w"))
    np.savez_compressed(os.path.join(OUT, "eval_attack_golden.npz"), **arrays)
    print("charmer cases", len(charmer), "bruteforce cases", len(brute))


def gen_hf_tokenizer():
    """transformers.CLIPTokenizer (slow, no ftfy) built from the reference's own BPE file, as eval_textfare.py's HF path
    uses it through tokenizer_wrapper (utils_attacks.py:67-71): padding=True, truncation=True."""
    import gzip
    import tempfile
    from open_clip.tokenizer import bytes_to_unicode, default_bpe
    from transformers import CLIPTokenizer
    merges = gzip.open(default_bpe()).read().decode("utf-8").split("\n")[1:49152 - 256 - 2 + 1]
    vocab = list(bytes_to_unicode().values())
    vocab = vocab + [v + "</w>" for v in vocab] + ["".join(m.split()) for m in merges] + ["<|startoftext|>", "<|endoftext|>"]
    d = tempfile.mkdtemp()
    json.dump({t: i for i, t in enumerate(vocab)}, open(os.path.join(d, "vocab.json"), "w"))
    open(os.path.join(d, "merges.txt"), "w").write("#version: 0.2\n" + "\n".join(merges) + "\n")
    tok = CLIPTokenizer(os.path.join(d, "vocab.json"), os.path.join(d, "merges.txt"), model_max_length=77)
    rng = random.Random(13)
    texts = ["a photo of a cat", "It's a DOG's life, isn't it?", "hello   world!!", "x" * 300, "a_b &amp; c &lt;", "3d 42 e.g. a cat.",
             "x <|endoftext|> y", "<|startoftext|>a", "x <end_of_text> y", "", " ", "a\tb\nc", "don't'll", "#$%&'()*+", "a" * 90 + " " + "zq " * 80]
    alpha = string.ascii_lowercase * 3 + string.ascii_uppercase + string.digits + "   '&;<_#x|" + string.punctuation
    for _ in range(400):
        texts.append("".join(rng.choice(alpha) for _ in range(rng.randint(1, 70))))
    caps = synth.make_captions(16, seed=3, kind="typical") + synth.make_captions(3, seed=3, kind="dense-77")
    for S in caps:
        texts.append(S)
        for _ in range(6):
            texts.append(utils_attacks.generate_sentence(S, rng.randint(0, 2 * len(S)), rng.randrange(len(V)), V, 1, alternative=-1))
    enc = [[t, tok(t).input_ids] for t in texts]
    batches = [texts[:7], caps[:5], [caps[0], caps[17]], ["a", "x" * 300]]
    wrapped = [[b, utils_attacks.tokenizer_wrapper(tok)(b).tolist()] for b in batches]
    json.dump({"pad_id": tok.pad_token_id, "eos_id": tok.eos_token_id, "encode": enc, "wrapped": wrapped},
              open(os.path.join(OUT, "hf_tokenizer_golden.json"), "w"))
    print("hf tokenizer strings", len(enc), "pad", tok.pad_token_id)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    if sys.argv[1:] == ["convert_ids"]:                       # add this fixture without regenerating the others
        gen_convert_ids()
        sys.exit(0)
    if sys.argv[1:] == ["latin1"]:
        gen_tokenizer_latin1()
        gen_attack_latin()
        sys.exit(0)
    gen_convert_ids()
    gen_edit()
    gen_tokenizer()
    gen_tower()
    gen_attack()
    gen_eval_attacks()
    gen_hf_tokenizer()
    gen_tokenizer_latin1()
    gen_attack_latin()
