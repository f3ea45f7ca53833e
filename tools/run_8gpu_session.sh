#!/bin/bash
# One 8-GPU lease: C3 in the driver's command shape, C4 (LINDEP stress, N = 5e7), C5 (FEAST, one node per GPU).
#   gpurun --gpus 8 -- 'bash tools/run_8gpu_session.sh'
N=${1:-8}
mkdir -p gpurun_out
free -g | head -2; nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # name, timeout, args...
  local name=$1 to=$2; shift 2
  timeout $to $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/r2_${name}_${N}gpu.log 2> gpurun_out/r2_${name}_${N}gpu.err
  echo "$name rc=$?"; tail -1 gpurun_out/r2_${name}_${N}gpu.log | cut -c1-300; grep -E "bench_error|Error" gpurun_out/r2_${name}_${N}gpu.err | tail -3 | cut -c1-600
}
run c3 600 --steps 5 --warmup 3
run c4 600 --workload c4 --steps 1 --warmup 1
run c5 900 --workload c5 --steps 1 --warmup 1 --no-e2e --no-profile --feast-tasks
timeout 300 $TR --master-port 29911 tests/multirank_worker.py kernels onesided lindep > gpurun_out/r2_multirank_${N}.log 2>&1; echo "multirank rc=$?"; grep -E "PASS|FAIL" gpurun_out/r2_multirank_${N}.log | head -10
