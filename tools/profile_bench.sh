#!/bin/bash
# ncu evidence for bench.py (run under gpurun, one GPU).  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-davidson --no-cpu-baseline"
if [ "${SKIP_LIST:-0}" != "1" ]; then
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
fi
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:dgemm_tn_tma_kernel<\(int\)1, \(int\)8, \(int\)14' -s 1 -c 1 -o gpurun_out/prof_vvvv -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
