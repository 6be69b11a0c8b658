#!/bin/bash
# developer loop on the GPU box: bit-exactness + kernel time of the default NLM kernel, then one ncu capture of it
tag=${1:-x}
timeout 500 python tools/nlm_ab.py sym > gpurun_out/nlm_ab_$tag.jsonl 2>&1
for t in $NLM_TRY_THREADS; do FPB_NLM_THREADS=$t NLM_AB_SHORT=1 timeout 300 python tools/nlm_ab.py sym 2>&1 | sed "s/^/threads=$t /" >> gpurun_out/nlm_ab_$tag.jsonl; done
python - <<PY
import json
for l in open("gpurun_out/nlm_ab_$tag.jsonl"):
    pre = ""
    if l.startswith("threads="): pre, l = l.split(" ", 1)
    try: r = json.loads(l)
    except Exception: print(l[:800]); continue
    print(pre, r["kernel"], [c.get("mismatching_pixels", c.get("error", "")[:30]) for c in r["cases"]], r.get("all_runs_ms"))
PY
if [ "$2" != "noncu" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_nlm_sym -c 1 -o gpurun_out/nlm_sym_$tag python tools/nlm_only.py 296 > gpurun_out/ncu_sym.log 2>&1; tail -1 gpurun_out/ncu_sym.log
fi
