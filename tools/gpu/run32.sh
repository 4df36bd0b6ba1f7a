#!/bin/bash
# round-2 GPU call 32: 2-D persistent K1 with two CTAs of 6 warps per SM: parity of the 2-D shapes, C1 / C4 with 1 and 2 CTAs per SM
set -x
mkdir -p gpurun_out
B="--no-e2e-full --no-newton --no-cpu-baseline"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -x 2>&1 | tail -4 > gpurun_out/r2_tests32.log
cat gpurun_out/r2_tests32.log
for k in 1 2; do for c in C1 C4; do
AD4SM_K1_CTAS=$k python bench.py --config $c --steps 30 --warmup 5 $B > gpurun_out/r2_b32_${c}_ctas$k.json 2> gpurun_out/r2_b32_${c}_ctas$k.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_b32_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["parity_check"]["status"])
    except Exception as e:
        print(f, "ERR", e)
PY
