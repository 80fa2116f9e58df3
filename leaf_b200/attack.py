"""Drop-in mirror of the reference's LEAF attack (the hot path):

    attack_text_leaf(model, tokenizer, sentences, anchor_features, device, objective='l2', n=10, k=1, V=..., constrain=False,
                     debug=False) -> (Tensor[B,E], list[str])                /root/reference/utils_attacks.py:297-393
    attack_text = attack_text_leaf                                            /root/reference/utils_attacks.py:646-647

Same names, argument meaning, RNG consumption (global numpy RNG: B position draws, then B character draws, per
round - utils_attacks.py:317 then :357 -> :236) and return values. What changes is where the work happens:
candidate expansion + BPE, the text tower, the TextFARE score and the argmax all run on the B200 through
libleaf_b200.so; the host only draws the random numbers, ships captions + draws (a few KB) and rebuilds the B
winner strings from (z*, c*) with the edit rule.

`model` is a LeafTextTower (leaf_b200/tower.py), or any torch module / dict whose parameters follow open_clip's CLIP
naming or HF's CLIPTextModel naming (an engine is created and cached on first use). `tokenizer` is accepted for
signature compatibility and not used: token ids are produced on the device, bit-identical to SimpleTokenizer.
"""
from __future__ import annotations

import string

import numpy as np
import torch

from . import dist as D
from ._native import LeafError
from .engine import LeafEngine, OBJECTIVES, bind_module

# /root/reference/train_AT_text_only.py:93
V_DEFAULT = [-1] + [ord(c) for c in string.ascii_lowercase + " " + string.ascii_uppercase
                    + string.digits + string.punctuation]


def generate_sentence(S: str, z: int, c: int) -> str:
    """Closed form of generate_sentence(S, z, u, V, k=1, alternative=-1) with c = V[u]
    (utils_attacks.py:169-213): z even = slot before character z//2, z odd = character z//2; writing the
    placeholder ('_', -1, or the character already there) deletes / leaves the slot empty."""
    i = z // 2
    if z % 2 == 1:
        if c == -1 or chr(c) == S[i]:
            return S[:i] + S[i + 1:]
        return S[:i] + chr(c) + S[i + 1:]
    if c == -1 or chr(c) == "_":
        return S
    return S[:i] + chr(c) + S[i:]


def _to_dev(a: np.ndarray, dev) -> torch.Tensor:
    t = torch.from_numpy(a)
    return t.pin_memory().to(dev, non_blocking=True) if torch.device(dev).type == "cuda" else t


def _engine_of(model) -> LeafEngine:
    """The engine behind `model`: a LeafTextTower / LeafEngine (or an engine double with the same three methods), or any
    torch module in open_clip CLIP / HF CLIPTextModel(WithProjection) / HF CLIPModel naming, also behind a DDP-style
    `.module` holder (engine.bind_module: created on first use, cached on the module, re-cast when its parameters have
    changed in place since the last call)."""
    if isinstance(model, LeafEngine) or all(hasattr(model, a) for a in ("expand_tokenize", "encode_tokens", "score")):
        return model
    eng = getattr(model, "leaf_engine", None)
    if isinstance(eng, LeafEngine) and not isinstance(model, torch.nn.Module):
        return eng
    if isinstance(model, torch.nn.Module):
        return bind_module(model)
    raise LeafError("model must be a LeafTextTower, a LeafEngine, or a torch module in open_clip / HF CLIP naming")


def _fast_draw(N, n):
    """What np.random.choice(range(N), size=n, replace=n > N) returns and consumes (legacy RandomState: a prefix of
    permutation(N) without replacement, randint with), minus choice()'s per-call overhead (26 -> 6 us; the 2 B draws of a
    round are the whole host cost of attack_text_leaf)."""
    return np.random.permutation(N)[:n] if n <= N else np.random.randint(0, N, size=n)


def _fast_draw_matches_choice():
    st = np.random.get_state()
    try:
        for N, n in ((7, 3), (161, 50), (96, 50), (5, 9), (1, 1)):
            np.random.seed(12345)
            a, ra = np.random.choice(range(N), size=n, replace=n > N), np.random.random()
            np.random.seed(12345)
            b, rb = _fast_draw(N, n), np.random.random()
            if not (a.dtype == b.dtype and np.array_equal(a, b) and ra == rb):
                return False
        return True
    except Exception:
        return False
    finally:
        np.random.set_state(st)


# the shortcut is used only if THIS numpy draws the same numbers and leaves the global stream in the same state
_DRAW = _fast_draw if _fast_draw_matches_choice() else (lambda N, n: np.random.choice(range(N), size=n, replace=n > N))


def _valid_mask(constrain, sentences, SS, B, n, device):
    """utils_attacks.py:321-325 / :360-364: candidates failing the constraint are replaced by the current sentence.
    constrain=True runs the filter on the device (engine.load_words + leaf_constrain_mask); a callable
    constrain(sentences, SS) -> bool[B][n] (e.g. the reference's own valid_sentence_batched, :110-143, which needs the
    NLTK corpora) is evaluated on the host from the candidate strings."""
    valid = np.asarray(constrain(sentences, SS), dtype=np.uint8).reshape(B, n)
    return _to_dev(valid, device)


def attack_text_leaf(model, tokenizer, sentences, anchor_features, device=None, objective="l2", n=10, k=1, V=V_DEFAULT,
                     constrain=False, debug=False, *, shard=None, group=None):
    """utils_attacks.py:297-393. Extra keyword-only arguments (not in the reference):
    shard = None         every process attacks the batch it was given (the reference's data-parallel meaning);
    shard = "samples"    `sentences`/`anchor_features` are the GLOBAL batch, identical on every rank; rank r attacks its
                         slice of samples, winners are all-gathered, every rank returns the global result;
    shard = "candidates" global batch on every rank; rank r scores its slice of the n candidates of every sample and
                         the per-sample argmax is completed across ranks (leaf_b200/dist.py)."""
    if objective not in OBJECTIVES:
        raise ValueError(f"unknown objective {objective!r}")                  # reference: falls through with loss undefined
    if shard not in (None, "samples", "candidates"):
        raise ValueError(f"unknown shard mode {shard!r}")
    eng = _engine_of(model)
    on_device = constrain is True
    if on_device and not getattr(eng, "has_words", False):
        raise LeafError("constrain=True needs the word list of the reference's filter (NLTK data, not reproducible offline): "
                        "call engine.load_words(nltk.corpus.words.words()) once, or pass "
                        "constrain=<callable(sentences, SS) -> bool[B][n]> (e.g. utils_attacks.valid_sentence_batched)")
    valid_fn = constrain if callable(constrain) else None
    sentences = list(sentences)
    B = len(sentences)
    dev = eng.device
    V = list(V)
    Vt = np.asarray(V, dtype=np.int32)
    if Vt.min() < -1 or Vt.max() > 0x7F:
        raise LeafError("attack alphabet V must hold -1 or ASCII code points")
    rank, G = D.world(group) if shard else (0, 1)
    blo, bhi = D.shard_range(B, rank, G) if shard == "samples" else (0, B)      # samples this rank attacks
    jlo, jhi = D.shard_range(n, rank, G) if shard == "candidates" else (0, n)   # candidates this rank scores
    Bl, nl = bhi - blo, jhi - jlo
    anchor = anchor_features
    if objective in ("dissim", "sim"):
        anchor /= anchor.norm(dim=-1, keepdim=True)                           # in place, as :304-308
    anchor = anchor.to(device=dev, dtype=torch.float32)[blo:bhi].contiguous()
    normalize = objective in ("sim", "dissim")
    best_feat = None
    if Bl > 0 and nl > 0:
        eng.reserve(Bl * nl + Bl)
    for _ in range(k):
        lens = [len(S) for S in sentences]
        # --- host: the reference's draws for the WHOLE batch, in the reference's order (pre-drawn: SURVEY.md app. E) ---
        # The character draws (:236) follow the position draws (:317) in the global stream and nothing else consumes it in
        # between, so they are made AFTER phase 1 has been queued: the host draws while the device scores.
        positions = np.stack([_DRAW(2 * L + 1, n) for L in lens])                                                # :317
        mine = sentences[blo:bhi]
        pos_l = np.ascontiguousarray(positions[blo:bhi, jlo:jhi]).astype(np.int32)
        picks = np.zeros((2, B), dtype=np.int64)                              # global candidate index per phase
        feat_l = torch.zeros((Bl, eng.embed_dim), dtype=torch.float32, device=dev)
        ok2 = np.ones(B, dtype=bool)
        if Bl > 0 and nl > 0:
            caps_d, off_d = eng.upload_captions(mine)
            host_d = _to_dev(np.concatenate([pos_l.ravel(), np.full(Bl * nl, 32, dtype=np.int32)]), dev)
            pos_d, chr1_d = host_d[:Bl * nl], host_d[Bl * nl:]
            # --- phase 1: choose the position (a space at each drawn z), :316-353 ---
            valid1 = eng.constrain_mask(caps_d, off_d, Bl, nl, pos_d, chr1_d) if on_device else None     # :321-325
            if valid_fn is not None:
                SS = [[generate_sentence(S, int(z), 32) for z in pos_l[i]] for i, S in enumerate(mine)]
                valid1 = _valid_mask(valid_fn, mine, SS, Bl, nl, dev)
            tok, ln, base = eng.expand_tokenize(caps_d, off_d, Bl, nl, pos=pos_d, chr_=chr1_d, valid=valid1)
            feats = eng.encode_tokens(tok, ln, normalize, base, (Bl * nl, nl), trim=True)   # rows [0, Bl*nl): candidates, then Bl captions
            best1, _, loss1 = eng.score(feats, anchor, Bl, nl, objective, want_loss=(debug or shard == "candidates"))
        us = np.stack([_DRAW(len(V), n) for _ in sentences])                                                     # :236
        chars2 = Vt[us]
        chr2_l = np.ascontiguousarray(chars2[blo:bhi, jlo:jhi])
        if Bl > 0 and nl > 0:
            chr2_d = _to_dev(chr2_l.ravel(), dev)
        if shard == "candidates":
            if nl > 0:
                val1, idx1 = loss1.gather(1, best1.long().view(-1, 1)).squeeze(1), best1.long() + jlo
            else:                                                             # n < world size: this rank scores nothing
                val1 = torch.full((B,), float("-inf"), device=dev)
                idx1 = torch.full((B,), D.NO_CANDIDATE, dtype=torch.long, device=dev)
            _, g1 = D.cross_shard_argmax(val1, idx1, group)                   # global index of the best position
            zstar = torch.from_numpy(positions.astype(np.int32)).to(dev).gather(1, g1.view(-1, 1)).squeeze(1)
            pos2_d = zstar.view(-1, 1).expand(Bl, nl).contiguous()            # same position for every candidate
            sel = None
        elif Bl > 0:
            g1, pos2_d, sel = best1.long(), pos_d, best1
        if Bl > 0 and nl > 0:
            # --- phase 2: choose the character at the best position, :355-389 ---
            valid2 = eng.constrain_mask(caps_d, off_d, Bl, nl, pos2_d, chr2_d, sel) if on_device else None   # :360-364
            if valid_fn is not None:
                zs_l = positions[np.arange(blo, bhi), g1.cpu().numpy()]
                SS = [[generate_sentence(S, int(zs_l[i]), int(c)) for c in chr2_l[i]] for i, S in enumerate(mine)]
                valid2 = _valid_mask(valid_fn, mine, SS, Bl, nl, dev)
            tok, ln, base = eng.expand_tokenize(caps_d, off_d, Bl, nl, pos=pos2_d, chr_=chr2_d, sel=sel, valid=valid2)
            feats = eng.encode_tokens(tok, ln, normalize, base, (Bl * nl, nl), trim=True)
            best2, feat_l, loss2 = eng.score(feats, anchor, Bl, nl, objective, want_loss=(debug or shard == "candidates"))
            g2 = best2.long()
            okv = torch.ones(Bl, dtype=torch.bool, device=dev) if valid2 is None else \
                valid2.view(Bl, nl).gather(1, g2.view(-1, 1)).squeeze(1).bool()
        if shard == "candidates":
            if nl > 0:
                val2, idx2 = loss2.gather(1, g2.view(-1, 1)).squeeze(1), g2 + jlo
            else:
                val2 = torch.full((B,), float("-inf"), device=dev)
                idx2 = torch.full((B,), D.NO_CANDIDATE, dtype=torch.long, device=dev)
                okv = torch.zeros(B, dtype=torch.bool, device=dev)
            _, gg2 = D.cross_shard_argmax(val2, idx2, group)
            owner = torch.zeros(B, dtype=torch.long, device=dev)
            for r in range(G):
                lo, hi = D.shard_range(n, r, G)
                owner[(gg2 >= lo) & (gg2 < hi)] = r
            feat_l = D.broadcast_rows(feat_l, owner, group)                   # winner's features from the rank that has them
            okv = D.broadcast_rows(okv.to(torch.float32), owner, group) > 0.5
            g2 = gg2
        # --- one small D2H per round: the winner indices (+ tokenizer status) ---
        have = Bl > 0 and (nl > 0 or shard == "candidates")
        if shard == "samples" and G > 1:
            # ONE all-gather per round: every rank's winner features with its three small integers per sample riding along as
            # exactly representable floats (indices < 2^24) - [Bl, E + 3] fp32
            sizes = [D.shard_range(B, r, G)[1] - D.shard_range(B, r, G)[0] for r in range(G)]
            small = torch.stack([g1, g2, okv.long()], dim=1).to(torch.float32) if have else torch.zeros((Bl, 3), device=dev)
            both = D.all_gather_cat(torch.cat([feat_l, small], dim=1), sizes, group)
            best_feat = both[:, :eng.embed_dim].contiguous()
            allp = both[:, eng.embed_dim:].cpu().numpy().T.astype(np.int64)
            if have:
                eng.check_status()
        elif have:
            allp, best_feat = torch.stack([g1, g2, okv.long()]).cpu().numpy(), feat_l
            eng.check_status()
        else:
            allp, best_feat = np.zeros((3, 0), dtype=np.int64), feat_l
        picks[0], picks[1], ok2 = allp[0], allp[1], allp[2].astype(bool)
        zs = positions[np.arange(B), picks[0]]
        cs = chars2[np.arange(B), picks[1]]
        sentences = [generate_sentence(S, int(zs[i]), int(cs[i])) if ok2[i] else S for i, S in enumerate(sentences)]
        if debug:
            print("LEAF round: best positions", zs.tolist(), "chars", cs.tolist())
    return best_feat, sentences


def attack_text(*args, **kwargs):
    """utils_attacks.py:646-647."""
    return attack_text_leaf(*args, **kwargs)
