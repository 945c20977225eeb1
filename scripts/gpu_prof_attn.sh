#!/bin/bash
tag=${1:-r01b}
CMD="python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --recording-seconds 70"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_kernel -s 3 -c 1 -o gpurun_out/${tag}_attn -f $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
