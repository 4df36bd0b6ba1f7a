#!/bin/bash
# round-2 GPU call 24 (8-GPU box, final): strong scaling of the 1M mesh at 8 (push transport, exchange stream at high priority),
# weak scaling at 8, C5 = Hex8 Ogden 200^3 over 8 GPUs, strong at 8 through the in-library NCCL transport
set -x
mkdir -p gpurun_out
run() { # n tag args...
  n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2_${tag}.json 2> gpurun_out/r2_${tag}.err
}
run 8 strong24_8 --scaling strong --interface push
run 8 weak24_8 --scaling weak --interface push
run 8 c5_24_8 --config C5 --scaling weak --interface push
run 8 strong24_8_lib --scaling strong --interface lib
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_*24_8*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"], d["parity_check"]["status"])
    except Exception as e:
        print(f, "ERR", e)
PY
