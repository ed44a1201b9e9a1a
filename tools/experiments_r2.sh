#!/bin/bash
# One gpurun call that validates and times the default-off kernel variants prepared at the end of round 1
# (DESIGN.md §4/§8).  Usage:  gpurun --timeout 1200 -- 'bash tools/experiments_r2.sh'
# Outputs land in gpurun_out/: exp_tests.log, bench_<tag>.json, prof_*.ncu-rep (read them offline with
# `ncu -i … --page raw|source --csv`, summarise with tools/ncu_summary.py into profiles/).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/exp_gpu.txt 2>&1

echo "##### experimental variants: bit-identity + op-level timing"
VNB_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_experimental.py -x -q -s > gpurun_out/exp_tests.log 2>&1
echo "exit=$?"; tail -n 25 gpurun_out/exp_tests.log

summ() {  # one line per bench json
python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
r = d["roofline"]
print(sys.argv[1], round(d["value"]), "tok/s", round(d["ms_per_step"], 1), "ms", d["clocks"]["sm_mhz"], "MHz",
      "gemm", round(r["achieved"]), "TF/s", {k: round(v, 1) for k, v in r["breakdown_ms"].items() if v > 1})
PY
}
echo "##### bench per variant (default first; each ~40 s)"
i=0
for cfg in "VNB_NOOP=1" "VNB_PAIR_ARRIVE_CTA=1" "VNB_ATTN_P_TMEM=1" "VNB_ATTN_V2=1" "VNB_PAIR_ARRIVE_CTA=1 VNB_ATTN_V2=1"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 240 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err \
    && summ gpurun_out/bench_$tag.json || tail -n 5 gpurun_out/bench_$tag.err
  i=$((i+1))
done

echo "##### ncu: attention (both variants) and the residual GEMMs (register-staged and TMA epilogue), source-level"
for v in 0 1; do
  VNB_ATTN_P_TMEM=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 45 -c 1 -f \
    -o gpurun_out/prof_attn_ptmem$v python tools/profile_step.py > gpurun_out/ncu_attn_ptmem$v.log 2>&1
  echo "ncu attention attn_p_tmem=$v exit=$?"
done
VNB_ATTN_V2=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 45 -c 1 -f \
  -o gpurun_out/prof_attn_v2 python tools/profile_step.py > gpurun_out/ncu_attn_v2.log 2>&1
echo "ncu attention v2 exit=$?"
VNB_RESID_TMA=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 170 -c 4 -f \
  -o gpurun_out/prof_gemm_resid_tma python tools/profile_step.py > gpurun_out/ncu_gemm_resid_tma.log 2>&1
echo "ncu resid_tma exit=$?"
