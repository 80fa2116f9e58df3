#!/usr/bin/env python
"""LEAF candidates scored per second (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W]            # this repo's engine on N B200s (torchrun for N > 1)
    python bench.py --impl reference [--steps K --warmup W]    # the reference algorithm on the host CPU (oracle port)

One "step" = one attack_text_leaf call on one batch of synthetic captions: 2*k*B*rho candidates expanded, tokenized,
encoded by the text tower and scored. Prints ONE JSON line on rank 0.
  value : device-resident throughput (captions and draws already in HBM, CUDA events around the kernels)
  e2e   : the same through leaf_b200.attack_text_leaf with host strings (H2D of captions/draws and D2H of the
          winners inside the timed region)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from leaf_b200 import synth  # noqa: E402

METRIC = "leaf_candidates_scored_per_sec"
UNIT = "candidates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="ViT-H-14")
    ap.add_argument("--batch", type=int, default=128, help="captions per GPU per step")
    ap.add_argument("--rho", type=int, default=50)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--captions", default="typical", choices=["typical", "dense-77", "short"])
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the full FARE step (attack + K4 + AdamW) leg")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1415.6), d.get("hbm_gbs", 6452.2), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(a):
    return f"{a.model} text tower, LEAF k={a.k} rho={a.rho}, batch {a.batch} {a.captions} synthetic captions per GPU"


# ------------------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle port of attack_text_leaf on the host cores
# ------------------------------------------------------------------------------------------------------------------
def cpu_attack_rate(a, seconds, steps=1, warmup=0):
    """Times oracle.attack_text_leaf_oracle (fp32 torch CPU, all host threads) on a bounded sample of the workload:
    same tower shape, same caption generator, same rho and k, reduced batch so that one call fits `seconds`."""
    from oracle import leaf_oracle as O
    cfg = synth.TOWERS[a.model]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = synth.random_tower_state_dict(cfg, seed=0)
    frozen = synth.perturbed_copy(sd, seed=1, std=1e-3)
    otok = O.OracleTokenizer()
    enc = lambda t, normalize: O.encode_text(sd, t, cfg.heads, quick_gelu=cfg.quick_gelu, normalize=normalize)
    # calibrate: one 77-slot row batch to estimate seconds per candidate
    caps = synth.make_captions(max(a.batch, 4), seed=0, kind=a.captions)
    with torch.no_grad():
        t0 = time.perf_counter()
        enc(otok(caps[:2] * 4), False)
        per_cand = (time.perf_counter() - t0) / 8
    per_call_budget = seconds / max(steps + warmup, 1)
    B = int(max(1, min(a.batch, per_call_budget / (per_cand * 2 * a.k * a.rho))))
    caps = caps[:B]
    times = []
    with torch.no_grad():
        anchor = O.encode_text(frozen, otok(caps), cfg.heads, quick_gelu=cfg.quick_gelu)
        for it in range(warmup + steps):
            np.random.seed(it)
            t0 = time.perf_counter()
            O.attack_text_leaf_oracle(enc, otok, caps, anchor, objective="l2", n=a.rho, k=a.k)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    cands = 2 * a.k * B * a.rho
    ms = 1e3 * sum(times) / len(times)
    return dict(value=cands / (ms / 1e3), unit=UNIT, cores=threads, kind="port",
                sample=f"oracle port of attack_text_leaf (fp32 torch CPU, dense 77-slot rows as the reference computes), "
                       f"{a.model}, B={B} of {a.batch}, rho={a.rho}, k={a.k}, {a.captions} captions, {len(times)} call(s)"), ms, cands


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms, cands = cpu_attack_rate(a, seconds=120.0, steps=a.steps, warmup=min(a.warmup, 1))
    line = dict(metric=METRIC, value=base["value"], unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=workload_name(a), candidates_per_step=cands), cpu_baseline=base,
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return dict(sm_mhz=mhz[len(mhz) // 2] if mhz else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.samples))


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch.distributed as dist
    from leaf_b200 import attack_text_leaf
    from leaf_b200.attack import V_DEFAULT
    from leaf_b200.tower import LeafTextTower

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.TOWERS[a.model]
    B, n, k = a.batch, a.rho, a.k
    # the path shards by samples: every rank attacks its own batch, no collective on the data path (weak scaling)
    tower = LeafTextTower.random(a.model, seed=0, device=dev)
    eng = tower.leaf_engine
    caps = synth.make_captions(B, seed=100 + rank, kind=a.captions)
    frozen_sd = synth.perturbed_copy(tower.open_clip_state_dict(), seed=1, std=1e-3)
    frozen = LeafTextTower(frozen_sd, heads=cfg.heads, quick_gelu=cfg.quick_gelu, device=dev)
    anchor = frozen.encode_text(frozen.tokenizer(caps)).clone()
    del frozen, frozen_sd
    torch.cuda.empty_cache()
    eng.reserve(B * n + B)
    Vt = np.asarray(V_DEFAULT, dtype=np.int32)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def draws(seed):
        rs = np.random.RandomState(seed)
        pos = np.stack([rs.choice(range(2 * len(S) + 1), size=n, replace=n > 2 * len(S) + 1) for S in caps]).astype(np.int32)
        ch = Vt[np.stack([rs.choice(range(len(Vt)), size=n, replace=n > len(Vt)) for _ in caps])]
        return torch.from_numpy(pos).to(dev), torch.from_numpy(ch).to(dev)

    caps_d, off_d = eng.upload_captions(caps)
    space = torch.full((B * n,), 32, dtype=torch.int32, device=dev)
    flop_per_cand = []

    def device_step(pos_d, chr_d, record=None):
        """The hot path with inputs resident in HBM (k = 1 form: no host round trip between the phases)."""
        tok, ln, base = eng.expand_tokenize(caps_d, off_d, B, n, pos=pos_d, chr_=space)
        f = eng.encode_tokens(tok, ln, False, base, (B * n, n), trim=True)
        if record is not None:
            record.append((ln[:B * n], eng.last_rows()))
        best1, _, _ = eng.score(f, anchor, B, n, "l2")
        tok, ln, base = eng.expand_tokenize(caps_d, off_d, B, n, pos=pos_d, chr_=chr_d, sel=best1)
        f = eng.encode_tokens(tok, ln, False, base, (B * n, n), trim=True)
        if record is not None:
            record.append((ln[:B * n], eng.last_rows()))
        return eng.score(f, anchor, B, n, "l2")

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- device-resident timing ----
    all_draws = [draws(s) for s in range(a.warmup + a.steps)]
    for i in range(a.warmup):
        device_step(*all_draws[i])
    torch.cuda.synchronize()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    lens_rec = []
    eng.launch_count(reset=True)
    with Clocks(local) as clk:
        torch.cuda.synchronize()
        for i in range(a.steps):
            flush.fill_(i)                                  # L2 flush between timed iterations (outside the events)
            ev[i][0].record()
            for _ in range(k):
                device_step(*all_draws[a.warmup + i])
            ev[i][1].record()
        torch.cuda.synchronize()
        launches = eng.launch_count()
        barrier()
        dev_ms = sum(s.elapsed_time(e) for s, e in ev)
        # ---- end to end through the public API (host strings in, winners out) ----
        for i in range(min(a.warmup, 2)):
            np.random.seed(1000 + i)
            attack_text_leaf(tower, None, caps, anchor.clone(), dev, objective="l2", n=n, k=k)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            np.random.seed(2000 + i)
            feats, adv = attack_text_leaf(tower, None, caps, anchor.clone(), dev, objective="l2", n=n, k=k)
            float(feats[0, 0].item())                       # device -> host read of the step's result
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        # ---- the same with the reference's --constrain filter (every shipped training script enables it), masks computed
        #      on the device against a synthetic word list; the headline stays the unconstrained attack ----
        eng.load_words(synth.word_list(100 + rank))
        np.random.seed(1500)
        attack_text_leaf(tower, None, caps, anchor.clone(), dev, objective="l2", n=n, k=k, constrain=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(a.steps):
            np.random.seed(2000 + i)
            feats, adv_c = attack_text_leaf(tower, None, caps, anchor.clone(), dev, objective="l2", n=n, k=k, constrain=True)
            float(feats[0, 0].item())
        torch.cuda.synchronize()
        con_ms = (time.perf_counter() - t0) * 1e3
        changed = sum(x != y for x, y in zip(adv_c, caps))
        barrier()
    t = torch.tensor([dev_ms, e2e_ms, con_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, con_ms = t.tolist()
    cands_step = 2 * k * B * n
    total_cands = cands_step * a.steps * world

    # ---- the whole FARE training step (utils_AT.py:291-366): attack + winners' forward/backward (K4) + data-parallel
    #      gradient all-reduce (NCCL) + AdamW + refresh of the engine's bf16 operand copies. Reported next to the
    #      headline; the headline metric stays "candidates scored per second" of the attack itself. ----
    train = None
    if not a.no_train_step:
        from leaf_b200.fare import FareTrainer
        frozen2 = LeafTextTower(synth.perturbed_copy(tower.open_clip_state_dict(), seed=1, std=1e-3), heads=cfg.heads,
                                quick_gelu=cfg.quick_gelu, device=dev)
        trainer = FareTrainer(tower, frozen2, rho=n, k_adv=k, lr=1e-5, wd=1e-4, beta1=0.9, beta2=0.98, eps=1e-6)   # scripts/train_leaf_vith.sh

        def train_step(seed):
            np.random.seed(seed)
            loss, _ = trainer.step(caps)
            return float(loss.item())

        train_step(3000)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        losses = [train_step(3001 + i) for i in range(a.steps)]
        torch.cuda.synchronize()
        tr_ms = (time.perf_counter() - t0) * 1e3 / a.steps
        tt = torch.tensor([tr_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        train = dict(ms_per_step=tt.item(), attack_ms_per_step=e2e_ms / a.steps, loss_first=losses[0], loss_last=losses[-1],
                     what="leaf_b200.fare.FareTrainer.step: frozen anchors + attack_text_leaf + tokenize winners + forward/backward "
                          "(K4, gradients accumulated into one flat buffer) + " + ("NCCL all-reduce + " if world > 1 else "")
                          + "native AdamW (one launch) + weight refresh")
        tower.trainable(False)
        del trainer, frozen2

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): one extra step with CUDA events around every GEMM launch ----
    rec = []
    device_step(*all_draws[-1], record=rec)                  # row counts (synchronises) + warm-up after the legs above
    torch.cuda.synchronize()
    eng.set_timing(True)
    for i in range(a.steps):                                 # averaged over the same number of steps as the headline
        flush.fill_(i)
        device_step(*all_draws[-1])                          # the draws whose row counts were recorded above
    torch.cuda.synchronize()
    per_step = lambda t: (t[0] / a.steps, t[1] // a.steps)
    gemm_ms, gemm_launches = per_step(eng.gemm_time_ms())
    insitu = {name: round(per_step(eng.class_time_ms(i))[0], 3) for i, name in enumerate(("gemm", "layernorm", "attention", "pack_embed"))}
    for epi, name in enumerate(("gemm_qkv_bf16", "gemm_fc1_bf16_act", "gemm_out_fc2_residual", "gemm_proj_f32", "gemm_fc2_residual", "gemm_out_bf16")):
        ms_epi, n_epi = per_step(eng.class_time_ms(4 + epi))
        if n_epi:
            insitu[name] = [round(ms_epi, 3), n_epi]
    eng.set_timing(False)
    lens = torch.cat([r[0] for r in rec]).double().cpu().numpy()
    rows_exec = sum(r[1] for r in rec)                                        # packed rows the GEMMs really processed
    W, L, E = cfg.width, cfg.layers, cfg.embed_dim
    # executed by the GEMM launches: every packed row through L-1 whole layers and the final layer's QKV projection; the
    # final layer's out-proj + MLP and the text projection see one (pooled EOS) row per sequence
    seqs = len(rec) * (B * n + B)
    gemm_flops = rows_exec * ((L - 1) * 24.0 + 6.0) * W * W + seqs * (18.0 * W * W + 2.0 * W * E)
    gemm_flops_credit = float((L * 24.0 * lens * W * W + 2.0 * W * E).sum())   # GEMM share of F(t), SURVEY.md 8d
    alg_flops = float(sum(cfg.flops_for_length(int(x)) for x in lens))
    peak_tf, peak_gbs, peak_src = peaks()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    step_ms = dev_ms / a.steps
    traffic = None                                  # DRAM bytes per GEMM launch from the committed ncu --set full capture
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("workload") == workload_name(a):
            traffic = tj["dram_bytes_per_launch"]
    roofline = dict(bound="tensor", achieved=achieved, peak=peak_tf, unit="TFLOP/s", frac=achieved / peak_tf, traffic=traffic,
                    kernel="gemm2_bf16_tn_kernel (tcgen05.mma.cta_group::2, all four Linear layers + projection)", peak_source=peak_src, gemm_launches_per_step=gemm_launches,
                    gemm_ms_per_step=gemm_ms, gemm_share_of_step=gemm_ms / step_ms if step_ms else None,
                    in_situ_ms_per_step=insitu,
                    algorithmic_tflop_per_step=alg_flops / 1e12, executed_gemm_tflop_per_step=gemm_flops / 1e12,
                    credited_gemm_tflops=gemm_flops_credit / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None,
                    rows_executed_per_step=int(rows_exec), rows_without_prefix_sharing=int(lens.sum()),
                    whole_step_tflops=alg_flops / (step_ms * 1e-3) / 1e12 if step_ms else None,
                    dense77_equiv_tflops=cands_step * cfg.dense_flops_per_candidate / (step_ms * 1e-3) / 1e12 if step_ms else None,
                    mean_len=float(lens.mean()))
    h2d = int(sum(len(c) for c in caps) + 32 + 4 * (B + 1) + 3 * 4 * B * n) * k
    d2h = (2 * 4 * B + 4 + 4) * k
    line = dict(metric=METRIC, value=total_cands / (dev_ms * 1e-3), unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup,
                ms_per_step=step_ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=workload_name(a), candidates_per_step_per_gpu=cands_step, tower=a.model, width=W, layers=L,
                            rho=n, k=k, batch_per_gpu=B, captions=a.captions, l2="flushed between timed steps (256 MB write); "
                            "per-phase activations ~%.1f GB" % (rows_exec / 2 * W * 14 / 1e9), parallelism=f"dp{world} sample-sharded"),
                e2e=dict(value=total_cands / (e2e_ms * 1e-3), unit=UNIT, ms_per_step=e2e_ms / a.steps, h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=d2h),
                e2e_constrained=dict(value=total_cands / (con_ms * 1e-3), unit=UNIT, ms_per_step=con_ms / a.steps,
                                     what="attack_text_leaf(constrain=True): validity masks from leaf_constrain_mask (device), "
                                          "synthetic word list", sentences_changed_last_step=int(changed)),
                gpu_launches=int(launches), clocks=clk.summary(), roofline=roofline)
    if train is not None:
        line["train_step"] = train
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            try:
                line["cpu_baseline"], _, _ = cpu_attack_rate(a, seconds=a.cpu_seconds)
            except Exception as ex:                          # the baseline must never take the bench line down
                line["cpu_baseline"] = dict(value=None, unit=UNIT, cores=os.cpu_count(), kind="port", sample=f"failed: {ex}")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
