N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist_peer_$N.log 2>&1; echo "dist peer rc=$?"; grep -E "PASS|FAIL|rror" gpurun_out/dist_peer_$N.log | tr '\n' ' ' | cut -c1-400; echo
timeout 300 $TR --master-port 29531 tools/feast_nodes_check.py > gpurun_out/feast_nodes_$N.log 2>&1; echo "feast rc=$?"; grep -E "PASS|FAIL|rror" gpurun_out/feast_nodes_$N.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29532 tools/feast_nodes_check.py > gpurun_out/feast_nodes_1.log 2>&1; echo "feast1 rc=$?"; grep -E "PASS|FAIL|rror" gpurun_out/feast_nodes_1.log | cut -c1-400
timeout 600 $TR --master-port 29515 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_c3_${N}gpu_v3.log 2>&1; echo "c3 rc=$?"; tail -1 gpurun_out/bench_c3_${N}gpu_v3.log | cut -c1-200
