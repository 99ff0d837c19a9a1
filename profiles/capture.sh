#!/bin/bash
# Run on the GPU box:  gpurun --timeout 900 -- 'bash profiles/capture.sh r01'
# Produces (under gpurun_out/, copy the summaries into profiles/ afterwards):
#   launches_<tag>.csv      every kernel launch of a short bench run with its device time
#   encode_<tag>.ncu-rep    one `--set full` capture of the dominant kernel (encode)
#   train_<tag>.ncu-rep     one `--set full` capture of train_loop_kernel (config-1 corpus, 1000 merges)
TAG=${1:-r01}
mkdir -p gpurun_out
set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    -k regex:"encode_kernel|quantize_kernel|merge_kernel|argmax_kernel|count_kernel|train_loop_kernel" --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-train --e2e-records 4096 > gpurun_out/launches_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 \
    -o gpurun_out/encode_${TAG} -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-train --e2e-records 2048 > gpurun_out/encode_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:train_loop_kernel -c 1 \
    -o gpurun_out/train_${TAG} -f \
    python profiles/train_prof.py 1000 > gpurun_out/train_${TAG}.log 2>&1
ls -la gpurun_out
