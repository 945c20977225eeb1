#!/bin/bash
# one --set full capture of the GEMM kernels of a short bench run (B200_PROFILING.md recipe) -> gpurun_out/<tag>_gemm.ncu-rep
tag=${1:-r01h}
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --recording-seconds 70"
mkdir -p gpurun_out
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 8 -c 6 -o gpurun_out/${tag}_gemm -f $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
