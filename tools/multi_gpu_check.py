#!/usr/bin/env python
"""Multi-GPU parity of the sharded attack modes over NCCL: under torchrun, every rank runs the unsharded
attack on the whole batch AND the sample- / candidate-sharded attacks; all three must select the same winners (bit-equal
scores: every candidate row is computed by the same kernels whatever rank holds it). BASELINE config 4's shape in small:
ViT-g-14 text tower width (a shallow copy), k = 2, candidate-sharded.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leaf_b200 import attack_text_leaf, synth  # noqa: E402
from leaf_b200.fare import FareTrainer  # noqa: E402
from leaf_b200.tower import LeafTextTower  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = synth.TowerCfg("ViT-g-14-4L", 1024, 4, 16, 1024)
sd = synth.random_tower_state_dict(cfg, seed=0, exact_numpy=False, device="cpu")
tower = LeafTextTower(sd, heads=cfg.heads, device=dev)
frozen = LeafTextTower(synth.perturbed_copy(sd, seed=1, std=1e-2), heads=cfg.heads, device=dev)
B, n, k = 12, 50, 2
caps = synth.make_captions(B, seed=5)
anchor = frozen.encode_text(frozen.tokenizer(caps))
res = {}
for mode in (None, "samples", "candidates"):
    np.random.seed(7)
    feats, adv = attack_text_leaf(tower, None, caps, anchor.clone(), dev, objective="l2", n=n, k=k, shard=mode)
    res[mode] = (feats, adv)
ok = True
for mode in ("samples", "candidates"):
    same = res[mode][1] == res[None][1]
    close = torch.equal(res[mode][0], res[None][0])
    print(f"rank {rank}: shard={mode}: winners equal {same}, features bit-equal {close}", flush=True)
    ok &= same and close
# data-parallel FARE step: each rank its own micro-batch. Gradients are averaged (a) by ONE blocking all-reduce of the flat
# buffer after the backward, (b) slice by slice on a side stream WHILE the backward runs (leaf_set_backward_hook). Both must
# leave identical parameters on every rank, and (a) and (b) must agree (same sums, possibly another order inside NCCL).
state0 = tower.flat_params.clone()
outs = {}
for overlap in (False, True):
    tower.flat_params.copy_(state0)
    tower.refresh()
    tr = FareTrainer(tower, frozen, rho=20, k_adv=1, lr=1e-4, accum_freq=2, overlap_allreduce=overlap)
    for mb in range(2):                                  # two micro-batches, one optimizer step
        np.random.seed(100 + rank + 10 * mb)
        loss, _ = tr.step(synth.make_captions(8, seed=50 + rank + 10 * mb))
    flat = tower.flat_params.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same_params = torch.equal(flat, ref)
    outs[overlap] = flat
    print(f"rank {rank}: FARE step (overlap_allreduce={overlap}) loss {loss.item():.4f}, optimizer steps {tr.opt_step}, parameters identical "
          f"across ranks after the update: {same_params}, moved: {not torch.equal(flat, state0)}", flush=True)
    ok &= same_params and tr.opt_step == 1 and not torch.equal(flat, state0)
    del tr
# The weight gradients of the 16-48-tile shapes are split-K sums and the bias / LayerNorm gradients atomic sums (run-to-run order),
# and the first AdamW step turns every gradient into +-lr whatever its size, so single near-zero gradients may move by a few
# percent of lr between ANY two runs; what must agree is the update as a whole.
d = outs[True] - outs[False]
u = outs[False] - state0
rel = (d.norm() / u.norm()).item()
frac = (d.abs() > 1e-2 * u.abs().max()).float().mean().item()
print(f"rank {rank}: overlapped vs blocking exchange: |update difference| / |update| = {rel:.3e}, max |difference| {d.abs().max().item():.3e} "
      f"(update size {u.abs().max().item():.3e}), elements differing by more than 1 % of it: {frac:.2e}", flush=True)
ok &= rel <= 1e-3 and frac <= 1e-4
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if t.item() == 1 else "FAIL", f"world={world}", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
