#!/bin/bash
# round-2 GPU call 17 (8-GPU box): split assembly + side-stream exchange at scale: strong scaling of the 1M mesh at 8 and 4,
# weak scaling at 8 (push transport), strong at 8 with the in-library NCCL transport
set -x
mkdir -p gpurun_out
run() { # n scaling interface tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600+$1)) bench.py --gpus $1 --steps 20 --warmup 5 --scaling $2 --interface $3 > gpurun_out/r2_${4}.json 2> gpurun_out/r2_${4}.err
}
run 8 strong push strong17_8
run 8 weak push weak17_8
run 4 strong push strong17_4
run 8 strong lib strong17_8_lib
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_*17_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"], d["parity_check"]["status"])
    except Exception as e:
        print(f, "ERR", e)
PY
