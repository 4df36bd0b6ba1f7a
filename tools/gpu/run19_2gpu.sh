#!/bin/bash
# round-2 GPU call 19 (2-GPU box): exchange stream at high priority: 2-GPU bitwise tests, strong scaling at 2 (push, lib)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -4 > gpurun_out/r2_tests19.log
cat gpurun_out/r2_tests19.log
for ifc in push lib; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --interface $ifc > gpurun_out/r2_strong19_2_$ifc.json 2> gpurun_out/r2_strong19_2_$ifc.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_*19_2*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"], d["parity_check"]["status"])
    except Exception as e:
        print(f, "ERR", e)
PY
