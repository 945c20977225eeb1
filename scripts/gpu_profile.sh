#!/bin/bash
# ncu evidence for the current build (B200_PROFILING.md recipe): launch list + one --set full capture per hot kernel.
# usage: scripts/gpu_profile.sh <tag>   -> gpurun_out/<tag>_launches.csv, <tag>_{attn,gemm,mem,split}.ncu-rep
# Every ncu pass runs only after the same command has exited 0 without ncu.
tag=${1:-r02}
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --recording-seconds 70 --skip-library"
mkdir -p gpurun_out
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
$CMD > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 3 -c 2 -o gpurun_out/${tag}_attn -f $CMD > gpurun_out/${tag}_ncu2.log 2>&1
$CMD > gpurun_out/${tag}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 8 -c 6 -o gpurun_out/${tag}_gemm -f $CMD > gpurun_out/${tag}_ncu3.log 2>&1
$CMD > gpurun_out/${tag}_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fbank_kernel|layernorm_kernel|decimate" -c 3 -o gpurun_out/${tag}_mem -f $CMD > gpurun_out/${tag}_ncu4.log 2>&1
$CMD > gpurun_out/${tag}_plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"attn_split_tc|gemm_pair_kernel<.*, 1>" -s 1 -c 5 -o gpurun_out/${tag}_split -f $CMD > gpurun_out/${tag}_ncu5.log 2>&1
ls -la gpurun_out/ | tail -20
tail -3 gpurun_out/${tag}_ncu2.log
