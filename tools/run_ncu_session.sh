#!/bin/bash
# ncu evidence on ONE GPU (each command first runs plain and must exit 0):
#   gpurun -- 'bash tools/run_ncu_session.sh'
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras"
$B > gpurun_out/r2_ncu_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench_c3.csv $B > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_c3.py > gpurun_out/r2_ncu_plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_spmv_dia2|k_orth_step' -s 60 -c 4 -f -o gpurun_out/r2_prof_c3 python tools/prof_c3.py > gpurun_out/r2_ncu_prof.log 2>&1
echo "full capture rc=$?"
python tools/prof_formats.py > gpurun_out/r2_ncu_plain_formats.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_spmv_sell|k_spmv_kron' -s 2 -c 4 -f -o gpurun_out/r2_prof_formats python tools/prof_formats.py > gpurun_out/r2_ncu_formats.log 2>&1
echo "formats capture rc=$?"
ls -la gpurun_out/*.ncu-rep
