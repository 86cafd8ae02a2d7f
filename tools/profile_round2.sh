#!/bin/bash
# Round-2 ncu evidence (run under gpurun, ONE GPU).  Numbers printed under ncu are never bench values.
#   1. launch list (gpu__time_duration.sum) of one ADC(3) build + matvecs at nocc=40 / nvirt=400
#   2. --set full captures of the new bandwidth kernels (pair matvec, exchange-ordered slabs, expand, fold)
#   3. --set full capture of the packed vvvv GEMM (dram traffic of the dominant kernel)
set -u
mkdir -p gpurun_out
A3="python bench.py --method adc3 --steps 2 --warmup 1 --no-davidson --verify-samples 0"
A2="python bench.py --steps 2 --warmup 1 --no-davidson --no-cpu-baseline --verify-samples 0"
$A3 > gpurun_out/p2_plain_adc3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r02_adc3_launches.csv $A3 > gpurun_out/p2_ncu_list.log 2>&1
echo "launch list rc=$?"
$A3 > gpurun_out/p2_plain_adc3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:pair_matvec_kernel|vvvv_exchange_slab_kernel|ovvv_exchange_slab_kernel|chunk_sum_kernel' -s 2 -c 8 \
    -o gpurun_out/r02_adc3_bandwidth_kernels -f $A3 > gpurun_out/p2_ncu_bw.log 2>&1
echo "bandwidth kernels rc=$?"
$A2 > gpurun_out/p2_plain_adc2x.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:expand_antisym_kernel|pack_antisym_virt_kernel|pack_antisym_occ_kernel|expand_occ_kernel' -s 4 -c 8 \
    -o gpurun_out/r02_expand_fold_kernels -f $A2 > gpurun_out/p2_ncu_ef.log 2>&1
echo "expand/fold rc=$?"
$A2 > gpurun_out/p2_plain_adc2xb.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:dgemm_tn_tma_kernel<\(int\)1, \(int\)8, \(int\)14' -s 1 -c 1 -o gpurun_out/r02_vvvv_gemm -f $A2 > gpurun_out/p2_ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
ls -la gpurun_out | grep -E "r02_|p2_"
