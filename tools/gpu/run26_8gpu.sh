#!/bin/bash
# round-2 GPU call 25 (8-GPU box): where does a strong-scaling step wait?  (AD4SM_DIST_PROF: events after K1 / exchange / interior / end)
set -x
mkdir -p gpurun_out
AD4SM_DIST_PROF=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 bench.py --gpus 8 --steps 20 --warmup 5 --scaling strong --interface push --no-parity-check > gpurun_out/r2_prof26_8.json 2> gpurun_out/r2_prof26_8.err
grep DISTPROF gpurun_out/r2_prof26_8.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_prof26_8.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["kernel_ms"])
PY
