N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist_peer.log 2>&1; echo "dist peer rc=$?"; grep -E "PASS|FAIL|rror" gpurun_out/dist_peer.log | head
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --workload c3mid --no-cpu > gpurun_out/bench_c3mid_${N}gpu_early.log 2>&1; echo "c3mid early rc=$?"; tail -1 gpurun_out/bench_c3mid_${N}gpu_early.log | cut -c1-200
EIGB200_PUSH_EARLY=0 timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 2 --warmup 1 --workload c3mid --no-cpu > gpurun_out/bench_c3mid_${N}gpu_late.log 2>&1; echo "c3mid late rc=$?"; tail -1 gpurun_out/bench_c3mid_${N}gpu_late.log | cut -c1-200
