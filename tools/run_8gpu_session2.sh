#!/bin/bash
# Second 8-GPU lease: C3 after the final kernel changes, FEAST run to convergence with dynamic task pulling.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2b_c3_${N}gpu.log 2> gpurun_out/r2b_c3_${N}gpu.err; echo "c3 rc=$?"; tail -1 gpurun_out/r2b_c3_${N}gpu.log | cut -c1-200
timeout 900 $TR --master-port 29612 bench.py --gpus $N --workload c5conv --distribute dynamic --steps 1 --warmup 0 --no-e2e --no-profile > gpurun_out/r2b_c5conv_${N}gpu.log 2> gpurun_out/r2b_c5conv_${N}gpu.err; echo "c5conv rc=$?"; tail -1 gpurun_out/r2b_c5conv_${N}gpu.log | cut -c1-300; grep -E "bench_error" gpurun_out/r2b_c5conv_${N}gpu.err | tail -2 | cut -c1-800
