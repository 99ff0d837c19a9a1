#!/bin/bash
# compute-sanitizer passes over the parity tests (small cases).  gpurun -- 'bash profiles/sanitize.sh [files...]'
mkdir -p gpurun_out
FILES=${@:-tests/test_gpu_encode.py tests/test_gpu_quantize.py tests/test_gpu_post.py tests/test_gpu_train.py}
SKIP='not full_size and not config1 and not many_records and not large_vocab and not wide_alphabet'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $FILES -x -q -k "$SKIP and not overflow and not random_text" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit=$?" >> gpurun_out/sanitize_$tool.log
  tail -4 gpurun_out/sanitize_$tool.log
done
