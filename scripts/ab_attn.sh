#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
  for cfg in 1/1 0/1 2/1 1/0 1/2; do
    IFS=/ read poly stag <<< "$cfg"
    tag=p${poly}_s${stag}_r${rep}
    ZK_ATTN_POLY=$poly ZK_ATTN_STAGGER=$stag timeout 300 python bench.py --steps 6 --warmup 3 --cpu-seconds 0 --skip-library \
      > gpurun_out/ab_attn_$tag.json 2> gpurun_out/ab_attn_$tag.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_attn_$tag.json").read().strip().splitlines()[-1])
    k = d["kernel_ms_per_step"]
    print("$tag", round(d["value"], 1), round(d["ms_per_step"], 1), d["recheck"]["windows_per_step"], d["clocks"]["sm_mhz"],
          {a: round(k[a], 1) for a in ("attention", "gemm_qkv", "gemm_fc1", "gemm_fc2")})
except Exception as e:
    print("$tag", "failed", e)
PY
  done
done
