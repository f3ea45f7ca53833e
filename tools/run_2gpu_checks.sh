N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for cfg in "0 1" "1 1" "0 0" "1 0"; do set -- $cfg
EIGB200_PUBLISH_LATE=$1 EIGB200_COOP=$2 timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --workload c3mid --no-cpu > gpurun_out/bench_c3mid_2gpu_late$1_coop$2.log 2>&1; echo "late=$1 coop=$2 rc=$?"; tail -1 gpurun_out/bench_c3mid_2gpu_late$1_coop$2.log | cut -c1-120
done
