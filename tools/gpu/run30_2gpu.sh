#!/bin/bash
# round-2 GPU call 30 (2-GPU box): per-step timeline of the in-library path vs dist.py (AD4SM_DIST_PROF)
set -x
mkdir -p gpurun_out
for ifc in lib push; do
AD4SM_DIST_PROF=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --interface $ifc --no-parity-check > gpurun_out/r2_s30_$ifc.json 2> gpurun_out/r2_s30_$ifc.err
grep DISTPROF gpurun_out/r2_s30_$ifc.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_s30_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
