import os, sys, time, torch
sys.path.insert(0, "/root/repo")
from leaf_b200 import synth
from leaf_b200.tower import LeafTextTower
tower = LeafTextTower.random("ViT-H-14", seed=0).trainable()
caps = synth.make_captions(128, seed=100)
tok = tower.tokenizer(caps)
with torch.no_grad():
    anchor = tower.encode_text(tok) + 0.01
def step(report=False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    f = tower.encode_text(tok)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    loss = torch.nn.functional.mse_loss(anchor, f, reduction="none").sum(-1).mean()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter(); torch.cuda.synchronize(); t5 = time.perf_counter()
    if report:
        print(f"fwd host {1e3*(t1-t0):.2f} ms, fwd total {1e3*(t2-t0):.2f}; bwd host {1e3*(t4-t3):.2f} ms, bwd total {1e3*(t5-t3):.2f}")
for i in range(4): step(i > 0)
