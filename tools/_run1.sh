mkdir -p gpurun_out
for poly in 0 4 8 12; do
  echo "### VNB_ATTN_POLY=$poly"
  VNB_ATTN_POLY=$poly VNB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -x -q -s -k "v2" > gpurun_out/exp2_tests_$poly.log 2>&1
  echo "exit=$?"; grep -E "attention attn|passed|failed|T=768|T=3072|Error|error" gpurun_out/exp2_tests_$poly.log | tail -8
done
VNB_ATTN_POLY=8 VNB_ATTN_V2=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 45 -c 1 -f -o gpurun_out/prof_attn_v3 python tools/profile_step.py > gpurun_out/ncu_attn_v3.log 2>&1
echo "ncu exit=$?"
