#!/bin/bash
# developer loop on the GPU box: bit-exactness + kernel time of the default NLM kernel, then one ncu capture of it
tag=${1:-x}
timeout 500 python tools/nlm_ab.py sym > gpurun_out/nlm_ab_$tag.jsonl 2>&1
python - <<PY
import json
for l in open("gpurun_out/nlm_ab_$tag.jsonl"):
    try: r = json.loads(l)
    except Exception: print(l[:800]); continue
    print(r["kernel"], [c.get("mismatching_pixels", c.get("error", "")[:30]) for c in r["cases"]], r.get("all_runs_ms"))
PY
if [ "$2" != "noncu" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_nlm_sym -c 1 -o gpurun_out/nlm_sym_$tag python tools/nlm_only.py 296 > gpurun_out/ncu_sym.log 2>&1; tail -1 gpurun_out/ncu_sym.log
fi
