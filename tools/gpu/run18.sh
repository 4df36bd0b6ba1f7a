#!/bin/bash
# round-2 GPU call 18: one-Gauss-point simplices through the persistent kernel (AD4SM_P1_PERSISTENT=1): Tet4 / Tria parity,
# C3 bench with and without; C2 bench of the current tree
set -x
mkdir -p gpurun_out
B="--no-e2e-full --no-newton --no-cpu-baseline"
AD4SM_P1_PERSISTENT=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -x -k "tet or tria or Tet or Tria or c3 or golden or gauss" 2>&1 | tail -6 > gpurun_out/r2_tests18.log
cat gpurun_out/r2_tests18.log
for k in 0 1; do
AD4SM_P1_PERSISTENT=$k python bench.py --config C3 --steps 20 --warmup 5 $B > gpurun_out/r2_bench18_C3_p$k.json 2> gpurun_out/r2_bench18_C3_p$k.err
done
python bench.py --steps 20 --warmup 5 $B > gpurun_out/r2_bench18_c2.json 2> gpurun_out/r2_bench18_c2.err
python - <<'PY'
import json
for f in ("C3_p0","C3_p1","c2"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2_bench18_{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["parity_check"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2_bench18_C3_p1.err
