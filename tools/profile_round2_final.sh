#!/bin/bash
# Round-2 closing evidence (one B200).  Numbers printed under ncu are never bench values.
#   1. ncu launch list of the default bench command (shares per kernel of the ADC(2)-x step)
#   2. --set full capture of the packed vvvv GEMM inside bench.py (roofline.traffic)
#   3. --set full captures of the two skinny ovvv GEMMs of the ADC(2) step (40 x 400 tile, N-major B)
#   4. phases of the headline Davidson run
set -u
mkdir -p gpurun_out
A2X="python bench.py --steps 2 --warmup 1 --no-davidson --no-cpu-baseline --verify-samples 0"
A2="python bench.py --method adc2 --steps 2 --warmup 1 --no-davidson"
if [ "${SKIP_LIST:-0}" != "1" ]; then
$A2X > gpurun_out/pf_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/r02_bench_launches.csv $A2X > gpurun_out/pf_ncu_list.log 2>&1
echo "launch list rc=$?"
fi
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:dgemm_tn_tma_kernel<\(int\)1, \(int\)8, \(int\)14' -s 1 -c 1 -o gpurun_out/r02_vvvv_gemm_final -f $A2X > gpurun_out/pf_ncu_gemm.log 2>&1
echo "vvvv capture rc=$?"
$A2 > gpurun_out/pf_plain_adc2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:dgemm_tn_tma_kernel<\(int\)1, \(int\)(10|8), \(int\)5, \(int\)(5|2), \(int\)(4|6), \(bool\)0, \(bool\)(0|1)>' -s 4 -c 4 \
    -o gpurun_out/r02_adc2_ovvv_gemms -f $A2 > gpurun_out/pf_ncu_adc2.log 2>&1
echo "adc2 gemm capture rc=$?"
if [ "${SKIP_DAVIDSON:-0}" != "1" ]; then
timeout 400 python tools/prof_davidson_headline.py > gpurun_out/prof_davidson_headline.txt 2>&1
echo "davidson profile rc=$?"
fi
ls -la gpurun_out | grep -E "r02_|pf_|prof_"
