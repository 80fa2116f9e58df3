"""Drop-in mirrors of the reference's single-sentence evaluation attacks, on the same kernels as attack_text_leaf
(SURVEY.md 8f item 3):

    attack_text_charmer_inference(model, tokenizer, sentence, anchor_features, device, objective, n, k, V, constrain,
                                  debug, batch_size, model_2, model_2_anchor_features) -> (str, int)
                                                                              /root/reference/utils_attacks.py:451-580
    attack_text_bruteforce(model, tokenizer, sentence, anchor_features, device, batch_size, objective, k, V, constrain,
                           debug) -> (str, int)                               /root/reference/utils_attacks.py:395-449

Same names, arguments and return values. The candidate lists are the reference's, in the reference's order:
position-major, every character of V per position (generate_all_sentences, :275-295). The reference's batching loop
never evaluates the LAST candidate of a list (`end = min((i+1)*bs, len-1)`, :422/:488/:543); that is part of its
results, so it is reproduced: scores are taken over the first len-1 candidates. `batch_size` is accepted and has no
other effect here (the whole list is one batch on the device). Ties: torch.topk leaves the order of equal scores
unspecified; here equal scores rank by ascending candidate index (what torch.argmax does for k = 1).

The host builds no candidate strings: positions and characters go to the device as int32 arrays, the K1 kernel
expands and tokenizes them, and only the top positions / the winner index come back.
"""
from __future__ import annotations

import numpy as np
import torch

from ._native import LeafError
from .attack import V_DEFAULT, _engine_of, _to_dev, generate_sentence
from .engine import OBJECTIVES


def _score_list(eng, eng2, sentence, pos, chr_, anchor, anchor2, objective, valid=None, device_filter=False):
    """Scores of the candidate list {(pos[g,j], chr_[g,j])} of ONE sentence. The list is laid out as `groups` samples of
    `per` candidates that all hold the same caption (per = len(V) when positions repeat per character), so the tower's
    shared-prefix reuse and in-group duplicate elimination apply. `pos` may already live on the device (the top-k
    output: no host round trip between the phases). Returns the [groups, per] loss tensor of each tower."""
    groups, per = pos.shape
    dev = eng.device
    pos_d = pos if torch.is_tensor(pos) else _to_dev(np.ascontiguousarray(pos, dtype=np.int32), dev)
    chr_d = _to_dev(np.ascontiguousarray(chr_, dtype=np.int32), dev)
    valid_d = None if valid is None else _to_dev(np.ascontiguousarray(valid, dtype=np.uint8), dev)
    norm = objective in ("sim", "dissim")
    # whole groups per pass, at most MAX_SEQS sequences at once (the workspace is sized for 77 rows per sequence: a brute
    # force over a 500-character caption is 96 000 candidates; the reference walks its list in batch_size pieces as well)
    gstep = max(1, MAX_SEQS // (per + 1))
    losses, losses2, valids = [], [], []
    for g0 in range(0, groups, gstep):
        g1 = min(groups, g0 + gstep)
        ng = g1 - g0
        caps_d, off_d = eng.upload_captions([sentence] * ng)
        p_c, c_c = pos_d[g0:g1].contiguous(), chr_d[g0:g1].contiguous()
        v_c = None if valid_d is None else valid_d[g0:g1].contiguous()
        if device_filter:                                                      # constrain=True: the mask never leaves the device
            v_c = eng.constrain_mask(caps_d, off_d, ng, per, p_c, c_c)
        eng.reserve(ng * per + ng)
        tok, ln, base = eng.expand_tokenize(caps_d, off_d, ng, per, pos=p_c, chr_=c_c, valid=v_c)
        for e, a, dst in ((eng, anchor, losses), (eng2, anchor2, losses2)):
            if e is None:
                continue
            e.reserve(ng * per + ng)
            feats = e.encode_tokens(tok, ln, norm, base, (ng * per, per), trim=True)
            _, _, loss = e.score(feats, a.expand(ng, -1).contiguous(), ng, per, objective, want_loss=True)
            dst.append(loss)
        valids.append(v_c)
    cat = lambda xs: None if not xs or xs[0] is None else (xs[0] if len(xs) == 1 else torch.cat(xs, dim=0))
    return [cat(losses), cat(losses2), cat(valids)]


MAX_SEQS = 16384


def _prep(model, model_2, anchor_features, model_2_anchor_features, objective, V):
    if objective not in OBJECTIVES:
        raise ValueError(f"unknown objective {objective!r}")
    eng = _engine_of(model)
    eng2 = _engine_of(model_2) if model_2 is not None else None
    Vt = np.asarray(list(V), dtype=np.int32)
    if Vt.min() < -1 or Vt.max() > 0x7F:
        raise LeafError("attack alphabet V must hold -1 or ASCII code points")
    if objective in ("dissim", "sim"):                                        # in place, as :464-471 / :399-403
        anchor_features /= anchor_features.norm(dim=-1, keepdim=True)
        if model_2 is not None:
            model_2_anchor_features /= model_2_anchor_features.norm(dim=-1, keepdim=True)
    a1 = anchor_features.to(device=eng.device, dtype=torch.float32).reshape(1, -1)
    a2 = None if eng2 is None else model_2_anchor_features.to(device=eng2.device, dtype=torch.float32).reshape(1, -1)
    return eng, eng2, Vt, a1, a2


def _device_filter(eng, constrain) -> bool:
    if constrain is True and not getattr(eng, "has_words", False):
        raise LeafError("constrain=True needs engine.load_words(<the reference's NLTK word list>) first, or pass "
                        "constrain=<callable(sentences, SS)>")
    return constrain is True


def _valid(constrain, sentence, pos, chr_):
    """constrain(sentences, SS) -> bool[1][len]: the reference's valid_sentence_batched (utils_attacks.py:110-143) needs
    NLTK corpora, so it enters as a caller-supplied callable, exactly as in attack_text_leaf."""
    if not callable(constrain):
        return None
    SS = [generate_sentence(sentence, int(z), int(c)) for z, c in zip(pos.ravel(), chr_.ravel())]
    return np.asarray(constrain([sentence], [SS]), dtype=np.uint8).reshape(pos.shape)


def attack_text_charmer_inference(model, tokenizer, sentence, anchor_features, device=None, objective="l2", n=10, k=1,
                                  V=V_DEFAULT, constrain=False, debug=False, batch_size=20 * 128, model_2=None,
                                  model_2_anchor_features=None):
    """utils_attacks.py:451-580. One sentence at a time; `n` is the number of positions kept after the probe."""
    eng, eng2, Vt, a1, a2 = _prep(model, model_2, anchor_features, model_2_anchor_features, objective, V)
    on_device = _device_filter(eng, constrain)
    nv = len(Vt)
    dist = 0
    for dist in range(k):
        L = len(sentence)
        n1 = 2 * L + 1
        if n1 - 1 < 1:
            raise ValueError("attack_text_charmer_inference: no candidate is evaluated for an empty sentence")
        # ---- probe: a space at every position (:476-517); the last position is never scored ----
        pos1 = np.arange(n1, dtype=np.int32).reshape(1, n1)
        chr1 = np.full((1, n1), 32, dtype=np.int32)
        l1, l1b, _ = _score_list(eng, eng2, sentence, pos1, chr1, a1, a2, objective, _valid(constrain, sentence, pos1, chr1), on_device)
        kk = min(n, n1 - 1)
        top, _ = eng.topk(l1, kk, m=n1 - 1, score_b=l1b)                       # :519
        # ---- every character of V at the kept positions (:524-575) ----
        chr2 = np.tile(Vt.reshape(1, nv), (kk, 1))
        pos2, valid2 = top.view(kk, 1).expand(kk, nv).contiguous(), None
        if callable(constrain):                                               # the mask needs the strings: one extra D2H
            pos2 = np.repeat(top.cpu().numpy().reshape(kk, 1), nv, axis=1)
            valid2 = _valid(constrain, sentence, pos2, chr2)
        l2, l2b, vd = _score_list(eng, eng2, sentence, pos2, chr2, a1, a2, objective, valid2, on_device)
        win, _ = eng.topk(l2, 1, m=kk * nv - 1, score_b=l2b)                  # :575, last candidate never scored
        okd = torch.ones(1, dtype=torch.int32, device=win.device) if vd is None else vd.reshape(-1)[win.long()].to(torch.int32)
        res = torch.cat([top, win, okd]).cpu().numpy()                        # the only device -> host read of the round
        eng.check_status()
        g = int(res[-2])
        z, c = int(res[g // nv]), int(Vt[g % nv])
        ok = bool(res[-1])
        sentence = generate_sentence(sentence, z, c) if ok else sentence
        if debug:
            print(sentence)
    return sentence, dist + 1


def attack_text_bruteforce(model, tokenizer, sentence, anchor_features, device=None, batch_size=20 * 128, objective="l2",
                           k=1, V=V_DEFAULT, constrain=False, debug=False):
    """utils_attacks.py:395-449: every position x every character of V, one round ('bruteforce for k=1'); only the 'l2'
    and 'dissim' objectives exist there (anything else leaves its loss undefined and raises)."""
    if objective not in ("l2", "dissim"):
        raise ValueError(f"attack_text_bruteforce supports objectives 'l2' and 'dissim', got {objective!r}")
    eng, _, Vt, a1, _ = _prep(model, None, anchor_features, None, objective, V)
    on_device = _device_filter(eng, constrain)
    nv = len(Vt)
    n1 = 2 * len(sentence) + 1
    pos = np.repeat(np.arange(n1, dtype=np.int32).reshape(n1, 1), nv, axis=1)
    chr_ = np.tile(Vt.reshape(1, nv), (n1, 1))
    valid = _valid(constrain, sentence, pos, chr_)
    loss, _, vd = _score_list(eng, None, sentence, pos, chr_, a1, None, objective, valid, on_device)
    win, _ = eng.topk(loss, 1, m=n1 * nv - 1)                                # :447, last candidate never scored
    g = int(win.item())
    eng.check_status()
    ok = True if vd is None else bool(vd.reshape(-1)[g].item())
    out = generate_sentence(sentence, int(pos.ravel()[g]), int(chr_.ravel()[g])) if ok else sentence
    if debug:
        print(out)
    return out, 1
