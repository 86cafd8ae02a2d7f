#!/bin/bash
# A/B of the GEMM pacing hint and of the stream-K tail policy on the packed vvvv shape (one B200):
# CUDA-event times first (no profiler), then DRAM bytes / L2 hit rate of one launch under ncu.
# The two environment switches exist only with tools/studies/gemm_pacing/pacing.patch applied to
# adcc_b200/csrc/dgemm.cu (the study's build); the shipped library has no pacing and always uses the
# tail rule that won (ragged round alone when it holds half a tile per CTA).
set -u
mkdir -p gpurun_out
SHAPE=${1:-780x79800x79800}
OUT=gpurun_out/ab_gemm_pace.txt
: > $OUT
for cfg in "0 101" "1 101" "1 50" "0 50"; do
  set -- $cfg
  echo "== ADCCB_GEMM_PACE=$1 ADCCB_GEMM_SK_MIN_PERCENT=$2" >> $OUT
  ADCCB_GEMM_PACE=$1 ADCCB_GEMM_SK_MIN_PERCENT=$2 timeout 300 python tools/bench_gemm.py $SHAPE 16000x16000x16000 >> $OUT 2>&1
done
M="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"
for cfg in "0 101" "1 101" "1 50"; do
  set -- $cfg
  echo "== ncu ADCCB_GEMM_PACE=$1 ADCCB_GEMM_SK_MIN_PERCENT=$2" >> $OUT
  ADCCB_GEMM_PACE=$1 ADCCB_GEMM_SK_MIN_PERCENT=$2 timeout 400 ncu --metrics $M --clock-control none \
    -k regex:dgemm_tn_tma --launch-skip 2 -c 1 --csv --log-file gpurun_out/ab_gemm_pace_ncu_$1_$2.csv \
    python tools/bench_gemm.py $SHAPE > gpurun_out/ab_gemm_pace_ncu_$1_$2.log 2>&1
  grep -v "^==PROF==" gpurun_out/ab_gemm_pace_ncu_$1_$2.csv | cut -d, -f5,13- >> $OUT
done
cat $OUT
