mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitizer_${tool}_r2.log 2>&1
  echo "$tool exit=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload done" gpurun_out/sanitizer_${tool}_r2.log | tail -3
done
timeout 300 python -m pytest tests/test_gpu_attention.py -x -q -s 2>&1 | grep -E "passed|failed|attention B=32|attention B=8" | tail -4
