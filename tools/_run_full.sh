mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_r2a.log 2>&1
echo "pytest exit=$?"; tail -n 6 gpurun_out/pytest_gpu_r2a.log; grep -E "^attention B=" gpurun_out/pytest_gpu_r2a.log | tail -16
timeout 400 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
echo "bench exit=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2a.json"))
r = d["roofline"]
print(round(d["value"]), "tok/s", round(d["ms_per_step"], 1), "ms", d["clocks"], "gemm", round(r["achieved"]), r["frac"])
print({k: round(v, 1) for k, v in r["breakdown_ms"].items()})
print(d.get("e2e"))
PY
VNB_ATTN_V2=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 45 -c 1 -f -o gpurun_out/prof_attn_r2 python tools/profile_step.py > gpurun_out/ncu_attn_r2.log 2>&1
echo "ncu exit=$?"
