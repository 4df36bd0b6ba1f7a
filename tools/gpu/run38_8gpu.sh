#!/bin/bash
# round-2 GPU call 38 (8-GPU box): final tree at 8 GPUs: strong and weak scaling (default transport)
set -x
mkdir -p gpurun_out
run() { n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2_${tag}.json 2> gpurun_out/r2_${tag}.err
}
run 8 strong38_8 --scaling strong
run 8 weak38_8 --scaling weak
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_*38_8*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"], d["parity_check"]["status"])
    except Exception as e:
        print(f, "ERR", e)
PY
