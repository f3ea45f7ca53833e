TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist_peer.log 2>&1; echo "dist peer rc=$?"; grep -E "PASS|FAIL|rror" gpurun_out/dist_peer.log | head
EIGB200_TRANSPORT=nccl timeout 300 $TR --master-port 29512 tools/dist_check.py > gpurun_out/dist_nccl.log 2>&1; echo "dist nccl rc=$?"; grep -E "PASS|FAIL" gpurun_out/dist_nccl.log | head
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 1 --workload c3mid > gpurun_out/bench_c3mid_2gpu_peer.log 2>&1; echo "c3mid peer rc=$?"; tail -1 gpurun_out/bench_c3mid_2gpu_peer.log | cut -c1-200
EIGB200_TRANSPORT=nccl timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 1 --warmup 1 --workload c3mid > gpurun_out/bench_c3mid_2gpu_nccl.log 2>&1; echo "c3mid nccl rc=$?"; tail -1 gpurun_out/bench_c3mid_2gpu_nccl.log | cut -c1-200
timeout 500 $TR --master-port 29515 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_c3_2gpu_peer.log 2>&1; echo "c3 peer rc=$?"; tail -1 gpurun_out/bench_c3_2gpu_peer.log | cut -c1-200
