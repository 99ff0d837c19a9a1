set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 -o gpurun_out/prof_encode_v4 python bench.py --steps 1 --warmup 3 --records 100000 --no-cpu --e2e-records 2048 > gpurun_out/prof4.log 2>&1
tail -3 gpurun_out/prof4.log
ls -la gpurun_out
