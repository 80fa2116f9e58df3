"""Deterministic synthetic inputs for the LEAF attack path.

* captions: "typical" (6-14 pseudo-words) and "dense-77" (40-60 pseudo-words, every row
  truncates to 77 tokens, /root/reference/src/open_clip/tokenizer.py:260-262).
* text-tower shapes of the four configs BASELINE.json names
  (/root/reference/src/open_clip/model_configs/ViT-{L,H,g,bigG}-14.json).
* random-init tower weights in the open_clip state-dict layout, following the init rule of
  /root/reference/src/open_clip/transformer.py:731-752 (std per tensor kind).

No model vocabulary: names follow the reference (captions, candidates, towers).
"""
from __future__ import annotations

import random
import string
from dataclasses import dataclass

import numpy as np
import torch

# attack alphabet, /root/reference/train_AT_text_only.py:93 (V[0] = -1 means "delete")
V_DEFAULT = [-1] + [ord(c) for c in string.ascii_lowercase + " " + string.ascii_uppercase
                    + string.digits + string.punctuation]

CONTEXT_LENGTH = 77
VOCAB_SIZE = 49408


@dataclass(frozen=True)
class TowerCfg:
    name: str
    width: int
    layers: int
    heads: int
    embed_dim: int
    quick_gelu: bool = False
    context_length: int = CONTEXT_LENGTH
    vocab_size: int = VOCAB_SIZE

    @property
    def dense_flops_per_candidate(self) -> float:
        """F_dense of SURVEY.md section 8d: what the reference executes per 77-slot row."""
        W, L, E, T = self.width, self.layers, self.embed_dim, self.context_length
        return L * (24.0 * T * W * W + 4.0 * T * T * W) + 2.0 * W * E

    def flops_for_length(self, t: int) -> float:
        """F(t) of SURVEY.md section 8d: causal work for a row whose EOS sits at index t-1."""
        W, L, E = self.width, self.layers, self.embed_dim
        return L * (24.0 * t * W * W + 2.0 * t * (t + 1) * W) + 2.0 * W * E


TOWERS = {
    "ViT-L-14": TowerCfg("ViT-L-14", 768, 12, 12, 768),
    "ViT-H-14": TowerCfg("ViT-H-14", 1024, 24, 16, 1024),
    "ViT-g-14": TowerCfg("ViT-g-14", 1024, 24, 16, 1024),
    "ViT-bigG-14": TowerCfg("ViT-bigG-14", 1280, 32, 20, 1280),
    # small shapes for parity tests (head_dim stays 64 as in every CLIP text tower)
    "tiny": TowerCfg("tiny", 128, 2, 2, 64),
    "small": TowerCfg("small", 256, 3, 4, 256),
}


def _pseudo_vocab(seed: int, size: int = 4096):
    rng = random.Random(1000003 + seed)
    words = []
    for _ in range(size):
        n = rng.randint(2, 9)
        words.append("".join(rng.choice(string.ascii_lowercase) for _ in range(n)))
    return words


def word_list(seed: int = 0, fraction: float = 0.5):
    """Stand-in for nltk.corpus.words.words() on the synthetic captions (the real list is NLTK data): a deterministic
    `fraction` of the pseudo-vocabulary make_captions(seed=...) draws from, so that roughly that share of a caption's
    words counts as dictionary words for the --constrain filter."""
    vocab = sorted(set(_pseudo_vocab(seed)))
    rng = random.Random(7919 + seed)
    return sorted(rng.sample(vocab, int(len(vocab) * fraction)))


def make_captions(batch: int, seed: int = 0, kind: str = "typical"):
    """Printable-ASCII captions; no '&', '<', '_', ';' in the base text."""
    rng = random.Random(seed)
    vocab = _pseudo_vocab(seed)
    lo, hi = {"typical": (6, 14), "dense-77": (40, 60), "short": (1, 4)}[kind]
    caps = []
    for _ in range(batch):
        n = rng.randint(lo, hi)
        caps.append(" ".join(rng.choice(vocab) for _ in range(n)))
    return caps


def tower_param_shapes(cfg: TowerCfg):
    """(key, shape, init-std or tag) in open_clip's CLIP state-dict naming."""
    W, E = cfg.width, cfg.embed_dim
    proj_std = (W ** -0.5) * ((2 * cfg.layers) ** -0.5)
    attn_std = W ** -0.5
    fc_std = (2 * W) ** -0.5
    out = [("token_embedding.weight", (cfg.vocab_size, W), 0.02),
           ("positional_embedding", (cfg.context_length, W), 0.01)]
    for i in range(cfg.layers):
        p = f"transformer.resblocks.{i}."
        out += [(p + "ln_1.weight", (W,), "ln_w"), (p + "ln_1.bias", (W,), "ln_b"),
                (p + "attn.in_proj_weight", (3 * W, W), attn_std),
                (p + "attn.in_proj_bias", (3 * W,), "bias"),
                (p + "attn.out_proj.weight", (W, W), proj_std),
                (p + "attn.out_proj.bias", (W,), "bias"),
                (p + "ln_2.weight", (W,), "ln_w"), (p + "ln_2.bias", (W,), "ln_b"),
                (p + "mlp.c_fc.weight", (4 * W, W), fc_std),
                (p + "mlp.c_fc.bias", (4 * W,), "bias"),
                (p + "mlp.c_proj.weight", (W, 4 * W), proj_std),
                (p + "mlp.c_proj.bias", (W,), "bias")]
    out += [("ln_final.weight", (W,), "ln_w"), ("ln_final.bias", (W,), "ln_b"),
            ("text_projection", (W, E), W ** -0.5)]
    return out


def random_tower_state_dict(cfg: TowerCfg, seed: int = 0, device="cpu", exact_numpy: bool = False):
    """Random-init text tower in open_clip layout.

    Matrix std's follow transformer.py:731-752. Biases and LayerNorm affine terms get small
    random values (the reference leaves them at zeros/ones at init, but trained checkpoints
    do not, and parity tests must exercise those terms).
    exact_numpy=True draws with numpy's RandomState so small golden fixtures are reproducible
    bit-for-bit on any box; large towers draw with torch on `device` for speed.
    """
    sd = {}
    if exact_numpy:
        rs = np.random.RandomState(seed)
        for key, shape, tag in tower_param_shapes(cfg):
            if tag == "ln_w":
                a = 1.0 + 0.1 * rs.standard_normal(shape)
            elif tag == "ln_b" or tag == "bias":
                a = 0.02 * rs.standard_normal(shape)
            else:
                a = tag * rs.standard_normal(shape)
            sd[key] = torch.from_numpy(a.astype(np.float32)).to(device)
        return sd
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    for key, shape, tag in tower_param_shapes(cfg):
        z = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        if tag == "ln_w":
            sd[key] = 1.0 + 0.1 * z
        elif tag == "ln_b" or tag == "bias":
            sd[key] = 0.02 * z
        else:
            sd[key] = tag * z
    return sd


def perturbed_copy(sd, seed: int = 1, std: float = 1e-3, exact_numpy: bool = False):
    """Frozen tower = copy + N(0, std) noise on every >=2-D weight so that the
    clean TextFARE loss is not identically zero."""
    out = {}
    if exact_numpy:
        rs = np.random.RandomState(seed)
        for k, v in sd.items():
            if v.ndim >= 2:
                noise = torch.from_numpy((std * rs.standard_normal(tuple(v.shape))).astype(np.float32))
                out[k] = v + noise.to(v.device)
            else:
                out[k] = v.clone()
        return out
    for i, (k, v) in enumerate(sd.items()):
        if v.ndim >= 2:
            g = torch.Generator(device=v.device)
            g.manual_seed(seed * 7919 + i)
            out[k] = v + std * torch.randn(v.shape, generator=g, device=v.device, dtype=v.dtype)
        else:
            out[k] = v.clone()
    return out
