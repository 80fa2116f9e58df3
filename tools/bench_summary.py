#!/usr/bin/env python
"""One line per bench JSON file: headline, clocks and the in-situ split (for same-box A/B runs)."""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    i = d["roofline"]["in_situ_ms_per_step"]
    g = lambda k: i.get(k, [float("nan")])[0]
    out = g("gemm_out_bf16") if "gemm_out_bf16" in i else g("gemm_out_fc2_residual") - g("gemm_fc2_residual")
    print(f"{f.split('/')[-1]:28s} {d['value']:9.0f} cand/s {d['ms_per_step']:7.2f} ms  e2e {d['e2e']['value']:8.0f}  {d['clocks']['sm_mhz']} MHz  "
          f"gemm {i['gemm']:.1f} (qkv {g('gemm_qkv_bf16'):.1f} out {out:.1f} fc1 {g('gemm_fc1_bf16_act'):.1f} fc2 {g('gemm_fc2_residual'):.1f}) "
          f"ln {i['layernorm']:.1f} att {i['attention']:.1f}  frac {d['roofline']['frac']:.3f}")
