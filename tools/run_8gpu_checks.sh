N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist_peer_$N.log 2>&1; echo "dist peer rc=$?"; grep -E "PASS|FAIL|rror" gpurun_out/dist_peer_$N.log | tr '\n' ' ' | cut -c1-400; echo
timeout 600 $TR --master-port 29515 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_c3_${N}gpu_final.log 2>&1; echo "c3 rc=$?"; tail -1 gpurun_out/bench_c3_${N}gpu_final.log | cut -c1-200
