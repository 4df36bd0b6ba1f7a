#!/bin/bash
# round-2 GPU call 4 (2-GPU box): whole parity suite incl. 2-GPU tests (torch NCCL transports + in-library NCCL), C3 bench with the new P=1 kernel
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_tests4.log
python bench.py --config C3 --steps 10 --warmup 3 --no-e2e-full > gpurun_out/r2_bench4_C3.json 2> gpurun_out/r2_bench4_C3.err
AD4SM_K1_T1=1 AD4SM_SORTED=0 python bench.py --config C3 --steps 10 --warmup 3 --no-e2e-full --no-cpu-baseline --no-parity-check > gpurun_out/r2_bench4_C3_old.json 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --interface lib > gpurun_out/r2_strong_lib_2.json 2> gpurun_out/r2_strong_lib_2.err
tail -8 gpurun_out/r2_tests4.log
