import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from leaf_b200 import synth
from leaf_b200.attack import V_DEFAULT
from leaf_b200.tower import LeafTextTower
B, n = 128, 50
dev = torch.device("cuda", 0)
mode = sys.argv[1]
if mode.startswith("bench"):
    tower = LeafTextTower.random("ViT-H-14", seed=0, device=dev)
else:
    sd = synth.random_tower_state_dict(synth.TOWERS["ViT-H-14"], seed=0, device=dev)
    tower = LeafTextTower({k: v.clone() for k, v in sd.items()}, heads=16, device=dev)
eng = tower.leaf_engine
caps = synth.make_captions(B, seed=100, kind="typical")
Vt = np.asarray(V_DEFAULT, dtype=np.int32)
def draws(seed):
    rs = np.random.RandomState(seed)
    pos = np.stack([rs.choice(range(2 * len(S) + 1), size=n, replace=n > 2 * len(S) + 1) for S in caps]).astype(np.int32)
    ch = Vt[np.stack([rs.choice(range(len(Vt)), size=n, replace=False) for _ in caps])]
    return torch.from_numpy(pos).to(dev), torch.from_numpy(ch).to(dev)
d, o = eng.upload_captions(caps)
eng.reserve(B * n + B)
space = torch.full((B * n,), 32, dtype=torch.int32, device=dev)
if "anchor" in mode:
    frozen = LeafTextTower(synth.perturbed_copy(tower.open_clip_state_dict(), seed=1, std=1e-3), heads=16, device=dev)
    anchor = frozen.encode_text(frozen.tokenizer(caps)).clone()
    del frozen
else:
    anchor = torch.randn((B, 1024), device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
rows = []
def step(pos, ch, rec=False):
    tok, ln, base = eng.expand_tokenize(d, o, B, n, pos=pos, chr_=space)
    f = eng.encode_tokens(tok, ln, False, base, (B * n, n), trim=True)
    if rec: rows.append(eng.last_rows())
    best = eng.score(f, anchor, B, n, "l2")[0]
    tok, ln, base = eng.expand_tokenize(d, o, B, n, pos=pos, chr_=ch, sel=best)
    f = eng.encode_tokens(tok, ln, False, base, (B * n, n), trim=True)
    if rec: rows.append(eng.last_rows())
    return eng.score(f, anchor, B, n, "l2")
D = [draws(s) for s in range(8)]
step(*D[0], rec=True)
for i in range(3): step(*D[i])
torch.cuda.synchronize()
for variant in ("same draws", "varying draws", "varying + flush", "same draws"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(5):
        if "flush" in variant: flush.fill_(i)
        step(*(D[3 + i] if "varying" in variant else D[0]))
    torch.cuda.synchronize()
    print(f"{mode:14s} {variant:16s} {(time.perf_counter() - t0) / 5 * 1e3:7.2f} ms   rows(step0) {rows}")
