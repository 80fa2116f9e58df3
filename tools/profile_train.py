#!/usr/bin/env python
"""One K4 pass (train-mode forward + backward of B winners) between cudaProfilerStart/Stop for ncu."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leaf_b200 import synth  # noqa: E402
from leaf_b200.tower import LeafTextTower  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "ViT-H-14"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
tower = LeafTextTower.random(model, seed=0).trainable()
caps = synth.make_captions(B, seed=100)
tok = tower.tokenizer(caps)
with torch.no_grad():
    anchor = tower.encode_text(tok) + 0.01


def step():
    f = tower.encode_text(tok)
    torch.nn.functional.mse_loss(anchor, f, reduction="none").sum(-1).mean().backward()


step()
torch.cuda.synchronize()
t0 = time.perf_counter()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("K4 fwd+bwd ms", (time.perf_counter() - t0) * 1e3, "launches", tower.leaf_engine.launch_count())
