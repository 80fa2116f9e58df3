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

from ._native import LeafError
from .engine import LeafEngine, OBJECTIVES

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


def _engine_of(model) -> LeafEngine:
    if isinstance(model, LeafEngine):
        return model
    eng = getattr(model, "leaf_engine", None)
    if isinstance(eng, LeafEngine):
        return eng
    if isinstance(model, torch.nn.Module):                        # bind a foreign tower (open_clip CLIP / HF) once
        module = model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model
        params = {k: v.detach() for k, v in module.state_dict(keep_vars=True).items()}
        heads = None
        for path in ("transformer.resblocks.0.attn.num_heads", "text.transformer.resblocks.0.attn.num_heads",
                     "config.num_attention_heads", "text_model.config.num_attention_heads"):
            obj = module
            try:
                for part in path.split("."):
                    obj = obj[int(part)] if part.isdigit() else getattr(obj, part)
                heads = int(obj)
                break
            except (AttributeError, IndexError, KeyError, TypeError):
                continue
        if heads is None:
            raise LeafError("cannot infer the number of attention heads of the tower")
        quick = "QuickGELU" in repr(type(getattr(getattr(module, "transformer", None), "resblocks", [None])[0]).__name__) \
            or any(type(m).__name__ == "QuickGELU" for m in module.modules())
        eng = LeafEngine(params, heads=heads, quick_gelu=quick)
        try:
            model.leaf_engine = eng
        except Exception:
            pass
        return eng
    raise LeafError("model must be a LeafTextTower, a LeafEngine, or a torch module in open_clip / HF CLIP naming")


def _valid_mask(constrain, sentences, SS, B, n, device):
    """utils_attacks.py:321-325 / :360-364: candidates failing the constraint are replaced by the current sentence.
    The reference's filter (valid_sentence_batched, :110-143) needs NLTK corpora; here it is a host callable
    constrain(sentences, SS) -> bool[B][n] supplied by the caller (parity unpinned)."""
    valid = np.asarray(constrain(sentences, SS), dtype=np.uint8).reshape(B, n)
    return torch.from_numpy(valid).pin_memory().to(device, non_blocking=True)


def attack_text_leaf(model, tokenizer, sentences, anchor_features, device=None, objective="l2", n=10, k=1, V=V_DEFAULT,
                     constrain=False, debug=False):
    if objective not in OBJECTIVES:
        raise ValueError(f"unknown objective {objective!r}")                  # reference: falls through with loss undefined
    eng = _engine_of(model)
    if constrain is True:
        raise LeafError("constrain=True needs the reference's NLTK word list, which cannot be reproduced offline; pass "
                        "constrain=<callable(sentences, SS) -> bool[B][n]> (e.g. utils_attacks.valid_sentence_batched)")
    valid_fn = constrain if callable(constrain) else None
    sentences = list(sentences)
    B = len(sentences)
    dev = eng.device
    V = list(V)
    Vt = np.asarray(V, dtype=np.int32)
    if Vt.min() < -1 or Vt.max() > 0x7F:
        raise LeafError("attack alphabet V must hold -1 or ASCII code points")
    anchor = anchor_features
    if objective in ("dissim", "sim"):
        anchor /= anchor.norm(dim=-1, keepdim=True)                           # in place, as :304-308
    anchor = anchor.to(device=dev, dtype=torch.float32).contiguous()
    normalize = objective in ("sim", "dissim")
    eng.reserve(B * n + B)
    best_feat = None
    for _ in range(k):
        lens = [len(S) for S in sentences]
        # --- host: the reference's draws, in the reference's order (pre-drawn: SURVEY.md appendix E) ---
        positions = np.stack([np.random.choice(range(2 * L + 1), size=n, replace=n > 2 * L + 1) for L in lens])  # :317
        us = np.stack([np.random.choice(range(len(V)), size=n, replace=(n > len(V))) for _ in sentences])       # :236
        chars2 = Vt[us]
        caps_d, off_d = eng.upload_captions(sentences)
        host = np.concatenate([positions.astype(np.int32).ravel(), np.full(B * n, 32, dtype=np.int32), chars2.ravel()])
        host_d = torch.from_numpy(host).pin_memory().to(dev, non_blocking=True)
        pos_d, chr1_d, chr2_d = host_d[:B * n], host_d[B * n:2 * B * n], host_d[2 * B * n:]
        # --- phase 1: choose the position (a space at each drawn z), :316-353 ---
        valid1 = None
        if valid_fn is not None:
            SS = [[generate_sentence(S, int(z), 32) for z in positions[i]] for i, S in enumerate(sentences)]
            valid1 = _valid_mask(valid_fn, sentences, SS, B, n, dev)
        tok, ln, base = eng.expand_tokenize(caps_d, off_d, B, n, pos=pos_d, chr_=chr1_d, valid=valid1)
        feats = eng.encode_tokens(tok, ln, normalize, base)          # rows [0, B*n) candidates, then the B captions
        best1, _, loss1 = eng.score(feats, anchor, B, n, objective, want_loss=debug)
        # --- phase 2: choose the character at the best position, :355-389 ---
        valid2 = None
        if valid_fn is not None:
            b1 = best1.cpu().numpy()
            zs = positions[np.arange(B), b1]
            SS = [[generate_sentence(S, int(zs[i]), int(c)) for c in chars2[i]] for i, S in enumerate(sentences)]
            valid2 = _valid_mask(valid_fn, sentences, SS, B, n, dev)
        tok, ln, base = eng.expand_tokenize(caps_d, off_d, B, n, pos=pos_d, chr_=chr2_d, sel=best1, valid=valid2)
        feats = eng.encode_tokens(tok, ln, normalize, base)
        best2, best_feat, loss2 = eng.score(feats, anchor, B, n, objective, want_loss=debug)
        # --- one small D2H per round: the 2B winner indices (+ tokenizer status) ---
        picks = torch.stack([best1, best2]).cpu().numpy()
        eng.check_status()
        zs = positions[np.arange(B), picks[0]]
        cs = chars2[np.arange(B), picks[1]]
        new = []
        for i, S in enumerate(sentences):
            ok = valid2 is None or bool(valid2[i, picks[1][i]].item())
            new.append(generate_sentence(S, int(zs[i]), int(cs[i])) if ok else S)
        if debug:
            print("LEAF round: best positions", zs.tolist(), "chars", cs.tolist())
        sentences = new
    return best_feat, sentences


def attack_text(*args, **kwargs):
    """utils_attacks.py:646-647."""
    return attack_text_leaf(*args, **kwargs)
