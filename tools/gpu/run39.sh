#!/bin/bash
# round-2 GPU call 39: final tree: whole GPU suite, smoke, default bench + reference arm
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4 > $O/r2_tests39.log
cat $O/r2_tests39.log
python __graft_entry__.py smoke 2>&1 | tail -1
( time python bench.py > $O/r2_final_c2.json 2> $O/r2_final_c2.err ) 2> $O/r2_final_c2.time
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_final_reference.json 2> $O/r2_final_reference.err
cat $O/r2_final_c2.time
python - <<'PY'
import json
for f in ("c2","reference"):
    d=json.loads([l for l in open(f"gpurun_out/r2_final_{f}.json") if l.startswith("{")][-1])
    pc=d.get("parity_check") or {}
    print(f, d.get("value"), d.get("ms_per_step"), d.get("kernel_ms"), (d.get("e2e") or {}).get("value"), pc.get("status"), (d.get("roofline") or {}).get("frac"), (d.get("roofline_path") or {}).get("frac"), d.get("gpu_launches"))
PY
