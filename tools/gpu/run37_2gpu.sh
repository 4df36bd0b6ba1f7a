#!/bin/bash
# round-2 GPU call 37 (2-GPU box): HALO scatter + interface rows of r on the exchange stream: parity (single-GPU emulation,
# 2-GPU bitwise tests of every transport), strong / weak at 2 with both paths
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -6 > gpurun_out/r2_tests37.log
cat gpurun_out/r2_tests37.log
for ifc in push lib; do
AD4SM_DIST_PROF=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --interface $ifc > gpurun_out/r2_s37_$ifc.json 2> gpurun_out/r2_s37_$ifc.err
grep DISTPROF gpurun_out/r2_s37_$ifc.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_w37_push.json 2> gpurun_out/r2_w37_push.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_?37_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"], d["parity_check"]["status"])
    except Exception as e:
        print(f, "ERR", e)
PY
