#!/bin/bash
# Row-sharded checks + the driver's exact bench command on N GPUs of one box:
#   gpurun --gpus N -- 'bash tools/run_multigpu_checks.sh N [steps] [warmup] [workload]'
N=${1:-2}; STEPS=${2:-2}; WARM=${3:-1}; WL=${4:-c3}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/multirank_worker.py > gpurun_out/multirank_${N}.log 2>&1; echo "multirank rc=$?"
grep -E "PASS|FAIL|rror" gpurun_out/multirank_${N}.log | head -20
timeout 1200 $TR --master-port 29515 bench.py --gpus $N --steps $STEPS --warmup $WARM --workload $WL \
  > gpurun_out/bench_${WL}_${N}gpu.log 2> gpurun_out/bench_${WL}_${N}gpu.err; echo "bench $WL rc=$?"
tail -1 gpurun_out/bench_${WL}_${N}gpu.log | cut -c1-400; tail -5 gpurun_out/bench_${WL}_${N}gpu.err
