TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist_peer.log 2>&1; echo "dist peer rc=$?"; grep -E "PASS|FAIL|rror" gpurun_out/dist_peer.log | head
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --workload c3mid --no-cpu > gpurun_out/bench_c3mid_2gpu_fused2.log 2>&1; echo "c3mid peer rc=$?"; tail -1 gpurun_out/bench_c3mid_2gpu_fused2.log | cut -c1-200
