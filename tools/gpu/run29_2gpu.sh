#!/bin/bash
# round-2 GPU call 29 (2-GPU box): why is the in-library push path 0.08 ms slower per step than dist.py's?  NCCL channel count / stream priority
set -x
mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --interface lib --no-parity-check > gpurun_out/r2_s29_$tag.json 2> gpurun_out/r2_s29_$tag.err
}
run base X=1
run nch1 NCCL_MAX_NCHANNELS=1
run prio0 AD4SM_COPY_PRIO=0
run nch1prio0 NCCL_MAX_NCHANNELS=1 AD4SM_COPY_PRIO=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_s29_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
