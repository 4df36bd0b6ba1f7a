#!/bin/bash
# round-2 multi-GPU call (8-GPU box): 2-GPU bitwise test, strong scaling of the 1M mesh at 2/4/8, weak at 8
set -x
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_mgpu_tests.log
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 5 --scaling strong > gpurun_out/r2_strong_$n.json 2> gpurun_out/r2_strong_$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_weak_8.json 2> gpurun_out/r2_weak_8.err
cat gpurun_out/r2_mgpu_tests.log
