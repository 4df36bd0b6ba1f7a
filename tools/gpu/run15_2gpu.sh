#!/bin/bash
# round-2 GPU call 15 (2-GPU box): split assembly + side-stream exchange: parity tests (single-GPU emulation, 2-GPU bitwise),
# strong scaling at 2 GPUs with the push and the in-library transports
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -8 > gpurun_out/r2_tests15.log
cat gpurun_out/r2_tests15.log
for ifc in push lib; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --interface $ifc > gpurun_out/r2_strong15_2_$ifc.json 2> gpurun_out/r2_strong15_2_$ifc.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_weak15_2.json 2> gpurun_out/r2_weak15_2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_*15_2*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"], d["parity_check"]["status"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/r2_strong15_2_push.err
