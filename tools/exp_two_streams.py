"""Experiment: one attack step of the bench workload on ONE engine (one stream) against the same step split into two
sample-halves on TWO engines / two CUDA streams, to see how much of the memory-bound work (LayerNorm, attention) the
hardware overlaps with the other half's GEMMs. LEAF_GEMM_PAIRS (if the engine reads it) caps the GEMM grid.

    python tools/exp_two_streams.py [halves=2]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leaf_b200 import synth                                  # noqa: E402
from leaf_b200.attack import V_DEFAULT                       # noqa: E402
from leaf_b200.tower import LeafTextTower                    # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, n = 128, 50
dev = torch.device("cuda", 0)
sd = synth.random_tower_state_dict(synth.TOWERS["ViT-H-14"], seed=0, device=dev)
caps = synth.make_captions(B, seed=100, kind="typical")
rs = np.random.RandomState(0)
Vt = np.asarray(V_DEFAULT, dtype=np.int32)
pos = np.stack([rs.choice(range(2 * len(S) + 1), size=n, replace=n > 2 * len(S) + 1) for S in caps]).astype(np.int32)
ch = Vt[np.stack([rs.choice(range(len(Vt)), size=n, replace=False) for _ in caps])]


def make(nparts):
    parts = []
    for i in range(nparts):
        lo, hi = B * i // nparts, B * (i + 1) // nparts
        tower = LeafTextTower({k: v.clone() for k, v in sd.items()}, heads=16, device=dev)
        eng = tower.leaf_engine
        d, o = eng.upload_captions(caps[lo:hi])
        b = hi - lo
        eng.reserve(b * n + b)
        parts.append(dict(eng=eng, d=d, o=o, b=b, pos=torch.from_numpy(pos[lo:hi]).to(dev), ch=torch.from_numpy(ch[lo:hi]).to(dev),
                          space=torch.full((b * n,), 32, dtype=torch.int32, device=dev),
                          anchor=torch.randn((b, 1024), device=dev), stream=torch.cuda.Stream()))
    return parts


def step(parts):
    best = [None] * len(parts)
    for i, p in enumerate(parts):                            # phase 1 of every part, then phase 2: launches are asynchronous
        with torch.cuda.stream(p["stream"]):
            tok, ln, base = p["eng"].expand_tokenize(p["d"], p["o"], p["b"], n, pos=p["pos"], chr_=p["space"])
            f = p["eng"].encode_tokens(tok, ln, False, base, (p["b"] * n, n), trim=True)
            best[i] = p["eng"].score(f, p["anchor"], p["b"], n, "l2")[0]
    for i, p in enumerate(parts):
        with torch.cuda.stream(p["stream"]):
            tok, ln, base = p["eng"].expand_tokenize(p["d"], p["o"], p["b"], n, pos=p["pos"], chr_=p["ch"], sel=best[i])
            f = p["eng"].encode_tokens(tok, ln, False, base, (p["b"] * n, n), trim=True)
            p["eng"].score(f, p["anchor"], p["b"], n, "l2")


def bench(parts, reps=5):
    for _ in range(2):
        step(parts)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step(parts)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


one, two = make(1), make(H)
for rep in range(2):
    print(f"1 stream: {bench(one):7.2f} ms   {H} streams: {bench(two):7.2f} ms")
