#!/bin/bash
# round-2 GPU call 20: w1 kernel with the sorted positions prefetched into shared memory (cp.async): P = 1 parity, C3 bench
set -x
mkdir -p gpurun_out
B="--no-e2e-full --no-newton --no-cpu-baseline"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_multi_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -6 > gpurun_out/r2_tests20.log
cat gpurun_out/r2_tests20.log
python bench.py --config C3 --steps 20 --warmup 5 $B > gpurun_out/r2_bench20_C3.json 2> gpurun_out/r2_bench20_C3.err
python - <<'PY'
import json
for f in ("C3",):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2_bench20_{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["parity_check"])
    except Exception as e:
        print(f, "ERR", e)
PY
