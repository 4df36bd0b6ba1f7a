#!/bin/bash
# round-2 GPU call 36: Ogden with integer exponents by multiplication: parity (Ogden cases, C5 slab vs oracle), C5 bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_ogden.py -m gpu -q --tb=short -x -k "ogden or Ogden or c5 or cas" 2>&1 | tail -4 > gpurun_out/r2_tests36.log
cat gpurun_out/r2_tests36.log
python bench.py --config C5 --steps 20 --warmup 5 > gpurun_out/r2_final_C5.json 2> gpurun_out/r2_final_C5.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_final_C5.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["kernel_ms"], d["parity_check"], d["roofline"]["frac"])
PY
