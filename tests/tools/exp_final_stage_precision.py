#!/usr/bin/env python
"""How much of the TextFARE-loss error comes from the LAST stage alone (ln_final output and text_projection rounded to bf16)?
fp32 oracle tower on the GPU for one phase of ViT-L-14 candidates; the same with only the final LayerNorm output and the projection
rounded to bf16; loss relative error of the second against the first."""
import os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from leaf_b200 import synth
from oracle import leaf_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "ViT-L-14"
cfg = synth.TOWERS[name]
B, n = 32, 50
sd = {k: v.cuda() for k, v in synth.random_tower_state_dict(cfg, seed=0, device="cuda").items()}
frozen = synth.perturbed_copy(sd, seed=1, std=1e-3)
caps = synth.make_captions(B, seed=100)
otok = O.OracleTokenizer()
rs = np.random.RandomState(0)
pos = np.stack([rs.choice(range(2 * len(S) + 1), size=n, replace=False) for S in caps])
strings = [O.edit_sentence(S, int(z), 32) for b, S in enumerate(caps) for z in pos[b]]
tok = otok(strings).cuda()

def pooled_ln(sd_, tok_):
    """everything up to and including ln_final on the pooled row (oracle lines, fp32)"""
    sd2 = dict(sd_)
    W = sd2["ln_final.weight"].numel()
    sd2["text_projection"] = torch.eye(W, device="cuda")
    outs = []
    with torch.no_grad():
        for s in range(0, tok_.shape[0], 512):
            outs.append(O.encode_text_device(sd2, tok_[s:s + 512], cfg.heads))
    return torch.cat(outs)

with torch.no_grad():
    a = pooled_ln(frozen, otok(caps).cuda()) @ frozen["text_projection"]
    p = pooled_ln(sd, tok)
    P = sd["text_projection"]
    exact = p @ P
    rounded = p.to(torch.bfloat16).float() @ P.to(torch.bfloat16).float()
    only_p = p.to(torch.bfloat16).float() @ P
    loss = lambda f: ((f.view(B, n, -1) - a.view(B, 1, -1)) ** 2).sum(-1)
    for tag, f in (("pooled and projection bf16", rounded), ("pooled bf16 only", only_p)):
        rel = (loss(f) - loss(exact)).abs() / loss(exact)
        print(f"{name}: {tag}: loss rel err max {rel.max().item():.2e}  p99 {torch.quantile(rel.flatten(), 0.99).item():.2e}  mean {rel.mean().item():.2e}")
