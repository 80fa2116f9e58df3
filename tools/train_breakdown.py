#!/usr/bin/env python
"""Where the FARE step's time goes (utils_AT.py:291-366 on the engine): wall-clock per stage with a synchronize after
each. Usage: python tools/train_breakdown.py [model] [batch] [rho]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leaf_b200 import attack_text_leaf, synth  # noqa: E402
from leaf_b200.tower import LeafTextTower  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "ViT-H-14"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = int(sys.argv[3]) if len(sys.argv) > 3 else 50
tower = LeafTextTower.random(model, seed=0).trainable()
caps = synth.make_captions(B, seed=100)
with torch.no_grad():
    anchor = tower.encode_text(tower.tokenizer(caps)) + 0.01
from leaf_b200.fare import FareTrainer  # noqa: E402
trainer = FareTrainer(tower, tower, rho=n, k_adv=1)
acc = {}


def stage(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return r


def step(seed):
    np.random.seed(seed)
    with torch.no_grad():
        _, adv = stage("attack", lambda: attack_text_leaf(tower, None, caps, anchor.clone(), "cuda", n=n, k=1))
        tok = stage("tokenize winners", lambda: tower.tokenizer(adv))
    f = stage("forward (train)", lambda: tower.encode_text(tok))
    loss = stage("loss", lambda: torch.nn.functional.mse_loss(anchor, f, reduction="none").sum(-1).mean())
    stage("backward", loss.backward)
    stage("optimizer (AdamW + zero_grad + refresh)", trainer.optimizer_step)


step(0)
acc.clear()
K = 3
for i in range(K):
    step(1 + i)
tot = sum(acc.values())
for k, v in acc.items():
    print(f"{k:20s} {v / K:9.2f} ms  {v / tot:6.1%}")
print(f"{'total':20s} {tot / K:9.2f} ms")
