"""LeafTextTower: a CLIP text tower whose parameters live in torch (open_clip CLIP naming)
and whose forward runs on the native engine. It offers the two duck-typed members the reference's attack uses:

    tower.encode_text(tokens, normalize=False) -> Tensor[N,E]     /root/reference/src/open_clip/model.py:269-284
    tower.tokenizer(list[str]) -> LongTensor[N,77]                 /root/reference/src/open_clip/tokenizer.py:226-265
"""
from __future__ import annotations

import torch

from . import synth
from .engine import LeafEngine


class _EncodeTextTrain(torch.autograd.Function):
    """encode_text under autograd: forward = leaf_forward_train, backward = leaf_backward. The parameters are inputs of
    the Function so that autograd accumulates the returned gradients into their .grad, like any torch module."""

    @staticmethod
    def forward(ctx, tower, tok, *params):
        ctx.tower = tower
        ctx.needs = [p.requires_grad for p in params]
        return tower.leaf_engine.forward_train(tok)

    @staticmethod
    def backward(ctx, dfeat):
        tower = ctx.tower
        names = list(tower._names)
        grads = {k: torch.zeros_like(getattr(tower, tower._names[k]).data) for k in names}
        tower.leaf_engine.backward(dfeat, grads)
        return (None, None) + tuple(grads[k] if need else None for k, need in zip(names, ctx.needs))


class LeafTextTower(torch.nn.Module):
    def __init__(self, state_dict: dict, heads: int, quick_gelu: bool = False, device="cuda"):
        super().__init__()
        self._names = {}
        for k, v in state_dict.items():
            safe = k.replace(".", "__")
            self._names[k] = safe
            self.register_parameter(safe, torch.nn.Parameter(v.detach().to(device=device, dtype=torch.float32).contiguous(),
                                                             requires_grad=False))
        self.heads, self.quick_gelu = heads, quick_gelu
        self.leaf_engine = LeafEngine(self.open_clip_state_dict(), heads=heads, quick_gelu=quick_gelu)

    @classmethod
    def random(cls, name_or_cfg, seed: int = 0, device="cuda", exact_numpy: bool = False):
        """Random-init tower of a named shape (synth.TOWERS; init rule of transformer.py:731-752)."""
        cfg = synth.TOWERS[name_or_cfg] if isinstance(name_or_cfg, str) else name_or_cfg
        sd = synth.random_tower_state_dict(cfg, seed=seed, device="cpu" if exact_numpy else device, exact_numpy=exact_numpy)
        return cls(sd, heads=cfg.heads, quick_gelu=cfg.quick_gelu, device=device)

    def open_clip_state_dict(self) -> dict:
        return {k: getattr(self, safe).data for k, safe in self._names.items()}

    def refresh(self):
        """Call after the parameters changed (optimizer step)."""
        self.leaf_engine.refresh_weights()

    def tokenizer(self, texts):
        return self.leaf_engine.tokenize(texts)

    def trainable(self, on: bool = True):
        """Mark the tower's parameters as requiring gradients (the attacked tower in train_AT_text_only.py)."""
        for p in self.parameters():
            p.requires_grad_(on)
        return self

    def encode_text(self, text, normalize: bool = False):
        """model.py:269-284. Under torch.no_grad() (the attack, utils_AT.py:295) this is the inference path; with
        gradients enabled and trainable parameters (utils_AT.py:317-319) the forward keeps its activations and
        loss.backward() runs the engine's backward. Call refresh() after optimizer.step()."""
        params = [getattr(self, safe) for safe in self._names.values()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            f = _EncodeTextTrain.apply(self, text, *params)
            return torch.nn.functional.normalize(f, dim=-1) if normalize else f
        with torch.no_grad():
            return self.leaf_engine.encode_tokens(text, None, normalize)
