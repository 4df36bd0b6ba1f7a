#!/bin/bash
# round-2 GPU call 14: single-lane TMA issue in K1 (no proxy fence): parity suite + C2 / C4 / C5 bench + per-part cycles
set -x
mkdir -p gpurun_out
B="--no-e2e-full --no-newton --no-cpu-baseline"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -8 > gpurun_out/r2_tests14.log
cat gpurun_out/r2_tests14.log
python bench.py --steps 20 --warmup 5 $B > gpurun_out/r2_bench14_c2.json 2> gpurun_out/r2_bench14_c2.err
python bench.py --config C4 --steps 20 --warmup 5 $B > gpurun_out/r2_bench14_C4.json 2> gpurun_out/r2_bench14_C4.err
python bench.py --config C5 --steps 10 --warmup 3 $B > gpurun_out/r2_bench14_C5.json 2> gpurun_out/r2_bench14_C5.err
AD4SM_K1_PROF=1 timeout 300 python bench.py --steps 2 --warmup 1 $B --no-parity-check 2>&1 >/dev/null | grep K1PROF | sort | uniq -c | sort -rn | head -4 > gpurun_out/r2_k1prof14.txt
cat gpurun_out/r2_k1prof14.txt
python - <<'PY'
import json
for f in ("c2","C4","C5"):
    d=json.loads([l for l in open(f"gpurun_out/r2_bench14_{f}.json") if l.startswith("{")][-1])
    print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["parity_check"])
PY
