import sys, torch
sys.path.insert(0, "/root/repo")
from leaf_b200 import synth
from leaf_b200.tower import LeafTextTower
eng = LeafTextTower.random(synth.TowerCfg("ab", 1024, 1, 16, 1024), seed=0).leaf_engine
g = torch.Generator(device="cuda").manual_seed(0)
M = 131072
A = (torch.randn((M, 1024), generator=g, device="cuda") * 0.5).to(torch.bfloat16)
Wq = (torch.randn((3072, 1024), generator=g, device="cuda") * 0.03).to(torch.bfloat16)
bq = torch.randn((3072,), generator=g, device="cuda")
C = torch.empty((M, 3072), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    eng.gemm(A, Wq, bq, 0, 0, C)
    torch.matmul(A, Wq.T, out=C)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.gemm(A, Wq, bq, 0, 0, C)
torch.matmul(A, Wq.T, out=C)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
# timing without the profiler
import time
for name, fn in (("ours", lambda: eng.gemm(A, Wq, bq, 0, 0, C)), ("cublas", lambda: torch.matmul(A, Wq.T, out=C))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(name, f"{ms*1e3:.1f} us  {2*M*1024*3072/ms/1e9:.0f} TFLOP/s")
