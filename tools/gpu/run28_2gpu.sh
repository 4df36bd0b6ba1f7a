#!/bin/bash
# round-2 GPU call 28 (2-GPU box): push transport inside ad4sm_dist_attach: 2-GPU bitwise tests, strong scaling at 2 with
# --interface lib (now push) and with AD4SM_DIST_PUSH=0 (send/recv), weak at 2 (default transport)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -12 > gpurun_out/r2_tests28.log
cat gpurun_out/r2_tests28.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --interface lib > gpurun_out/r2_strong28_2_lib.json 2> gpurun_out/r2_strong28_2_lib.err
AD4SM_DIST_PUSH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --interface lib > gpurun_out/r2_strong28_2_lib_sendrecv.json 2> gpurun_out/r2_strong28_2_lib_sendrecv.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_*28_2*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"], d["parity_check"]["status"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2_strong28_2_lib.err
