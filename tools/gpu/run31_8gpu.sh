#!/bin/bash
# round-2 GPU call 31 (8-GPU box): the in-library (C ABI) partitioned path with its push transport: strong and weak at 8
set -x
mkdir -p gpurun_out
run() { n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r2_${tag}.json 2> gpurun_out/r2_${tag}.err
}
run 8 strong31_8_lib --scaling strong --interface lib
run 8 weak31_8_lib --scaling weak --interface lib
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_*31_8*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"], d["parity_check"]["status"])
    except Exception as e:
        print(f, "ERR", e)
PY
