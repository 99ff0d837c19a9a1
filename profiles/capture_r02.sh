#!/bin/bash
# Run on the GPU box:  gpurun --timeout 900 -- 'bash profiles/capture_r02.sh'
# Produces under gpurun_out/ (summaries are copied into profiles/ afterwards):
#   launches_r02.csv         every kernel launch of a short bench run with its device time (cold, serialised)
#   encode_r02.ncu-rep       `--set full` capture of the 5,000-merge encode kernel (bitmap trie)
#   encode2_r02.ncu-rep      `--set full` capture of the pair-table walker on the 10,000-merge table
#   quantize_r02.ncu-rep     `--set full` capture of quantize_kernel
mkdir -p gpurun_out
set -x
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    -k regex:"encode|quantize|csr_|merge_kernel|argmax_kernel|count_kernel|loop_kernel|apply_lists|compact_delta" \
    --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-config4 --e2e-records 4096 > gpurun_out/launches_r02.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 \
    -o gpurun_out/encode_r02 -f \
    env ECGB_ENCODE_V1=1 python bench.py --steps 1 --warmup 3 --no-cpu --no-train --no-config3 --no-i16 --e2e-records 2048 > gpurun_out/encode_r02.log 2>&1
# the headline kernel of the final build: the pair-table walker on the 5,000-merge table
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode2_kernel -s 3 -c 1 \
    -o gpurun_out/encode2_r02_final -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-train --no-config3 --no-i16 --e2e-records 2048 > gpurun_out/encode2_r02_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode2_kernel -s 3 -c 1 \
    -o gpurun_out/encode2_r02 -f \
    python profiles/encode_ab.py 100000 10000 f32 > gpurun_out/encode2_r02.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:quantize_kernel -s 2 -c 1 \
    -o gpurun_out/quantize_r02 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-train --no-config3 --no-i16 --e2e-records 2048 > gpurun_out/quantize_r02.log 2>&1
ls -la gpurun_out | tail -8
