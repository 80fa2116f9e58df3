#!/usr/bin/env python
"""Where the FARE step's time goes outside the attack: every stage of FareTrainer.step (leaf_b200/fare.py) run on its own
between stream synchronisations, wall clock, ViT-H-14, B = 128, rho = 50 (bench.py's train_step leg runs them back to back).

    python tools/profile_fare_step.py [steps]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leaf_b200 import synth  # noqa: E402
from leaf_b200.fare import FareTrainer  # noqa: E402
from leaf_b200.tower import LeafTextTower  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cfg = synth.TOWERS["ViT-H-14"]
dev = torch.device("cuda:0")
tower = LeafTextTower.random("ViT-H-14", seed=0).trainable()
frozen = LeafTextTower(synth.perturbed_copy(tower.open_clip_state_dict(), seed=1, std=1e-3), heads=cfg.heads, quick_gelu=cfg.quick_gelu, device=dev)
tr = FareTrainer(tower, frozen, rho=50, k_adv=1, lr=1e-5, wd=1e-4, beta1=0.9, beta2=0.98, eps=1e-6)
caps = synth.make_captions(128, seed=0)
acc = {}


def timed(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
    return out


for it in range(2 + steps):
    np.random.seed(it)
    if it == 2:
        acc.clear()
    anchors = timed("anchors (frozen tower, tokenize + encode)", lambda: tr.anchors(caps))
    adv = timed("attack_text_leaf", lambda: tr.attack(caps, anchors.clone()))
    tok, lens = timed("tokenize winners", lambda: tower.tokenizer(adv, with_lengths=True))
    feats = timed("encode_text (training forward)", lambda: tower.encode_text(tok, host_lengths=lens))
    loss = timed("loss", lambda: torch.nn.functional.mse_loss(anchors, feats, reduction="none").sum(dim=-1).mean())
    timed("backward", lambda: loss.backward())
    timed("optimizer_step (AdamW + zero_grad + refresh)", lambda: tr.optimizer_step())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    np.random.seed(it)
    tr.step(caps)
    torch.cuda.synchronize()
    acc.setdefault("FareTrainer.step, back to back", []).append((time.perf_counter() - t0) * 1e3)

tot = 0.0
for k, v in acc.items():
    m = float(np.median(v))
    if not k.startswith("FareTrainer"):
        tot += m
    print(f"{k:50s} {m:8.3f} ms")
print(f"{'sum of the stages':50s} {tot:8.3f} ms")
