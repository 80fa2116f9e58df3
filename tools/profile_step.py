#!/usr/bin/env python
"""One device-resident LEAF attack step (2 phases) of the bench workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off`. Usage: python tools/profile_step.py [model] [batch] [rho] [captions]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leaf_b200 import synth  # noqa: E402
from leaf_b200.attack import V_DEFAULT  # noqa: E402
from leaf_b200.tower import LeafTextTower  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "ViT-H-14"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = int(sys.argv[3]) if len(sys.argv) > 3 else 50
kind = sys.argv[4] if len(sys.argv) > 4 else "typical"
tower = LeafTextTower.random(model, seed=0)
eng = tower.leaf_engine
caps = synth.make_captions(B, seed=100, kind=kind)
anchor = tower.encode_text(tower.tokenizer(caps)) + 0.01
rs = np.random.RandomState(0)
pos = torch.from_numpy(np.stack([rs.choice(range(2 * len(S) + 1), size=n, replace=n > 2 * len(S) + 1) for S in caps]).astype(np.int32)).cuda()
ch = torch.from_numpy(np.asarray(V_DEFAULT, dtype=np.int32)[rs.randint(0, 96, size=(B, n))]).cuda()
space = torch.full((B * n,), 32, dtype=torch.int32, device="cuda")
d, o = eng.upload_captions(caps)


def step():
    tok, ln, base = eng.expand_tokenize(d, o, B, n, pos=pos, chr_=space)
    f = eng.encode_tokens(tok, ln, False, base, (B * n, n), trim=True)
    best1, _, _ = eng.score(f, anchor, B, n, "l2")
    tok, ln, base = eng.expand_tokenize(d, o, B, n, pos=pos, chr_=ch, sel=best1)
    f = eng.encode_tokens(tok, ln, False, base, (B * n, n), trim=True)
    return eng.score(f, anchor, B, n, "l2")


step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("rows", eng.last_rows(), "launches", eng.launch_count())
