#!/usr/bin/env python
"""Experiment: does replaying one attack step as a CUDA graph (no launch gaps) beat 358 stream launches? Same workload as bench.py."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leaf_b200 import synth
from leaf_b200.attack import V_DEFAULT
from leaf_b200.tower import LeafTextTower

B, n = 128, 50
tower = LeafTextTower.random("ViT-H-14", seed=0)
eng = tower.leaf_engine
caps = synth.make_captions(B, seed=100)
anchor = tower.encode_text(tower.tokenizer(caps)) + 0.01
rs = np.random.RandomState(0)
pos = torch.from_numpy(np.stack([rs.choice(range(2 * len(S) + 1), size=n, replace=False) for S in caps]).astype(np.int32)).cuda()
ch = torch.from_numpy(np.asarray(V_DEFAULT, dtype=np.int32)[rs.randint(0, 96, size=(B, n))]).cuda()
space = torch.full((B * n,), 32, dtype=torch.int32, device="cuda")
d, o = eng.upload_captions(caps)
eng.reserve(B * n + B)

def step():
    tok, ln, base = eng.expand_tokenize(d, o, B, n, pos=pos, chr_=space)
    f = eng.encode_tokens(tok, ln, False, base, (B * n, n), trim=True)
    best1, _, _ = eng.score(f, anchor, B, n, "l2")
    tok, ln, base = eng.expand_tokenize(d, o, B, n, pos=pos, chr_=ch, sel=best1)
    f = eng.encode_tokens(tok, ln, False, base, (B * n, n), trim=True)
    return eng.score(f, anchor, B, n, "l2")

for _ in range(3):
    step()
torch.cuda.synchronize()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(2):
        step()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        out = step()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=8):
    ts = []
    for i in range(reps):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts[2:]) / len(ts[2:])
for rep in range(3):
    print("stream launches %.2f ms   graph replay %.2f ms" % (timeit(step), timeit(g.replay)))
