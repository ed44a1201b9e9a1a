mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2b.log 2>&1
echo "pytest exit=$?"; tail -n 15 gpurun_out/pytest_gpu_r2b.log | cut -c1-300
grep -E "^attention B=|embedding projection|vs fp32 reference|vs bf16 oracle|masked positions have|token agreement" gpurun_out/pytest_gpu_r2b.log | cut -c1-260 | tail -40
for c in 2 1 3 4; do
  timeout 500 python bench.py --config $c $( [ $c != 2 ] && echo --no-cpu-baseline ) > gpurun_out/bench_r2b_cfg$c.json 2> gpurun_out/bench_r2b_cfg$c.err
  echo "bench cfg$c exit=$?"; python - $c <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_r2b_cfg{sys.argv[1]}.json"))
    r = d["roofline"]
    print(round(d["value"]), "tok/s", round(d["ms_per_step"], 1), "ms rtf", round(d["rtf"]), d["clocks"]["sm_mhz"], "MHz gemm", round(r["achieved"]), round(r["frac"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
    print({k: round(v, 1) for k, v in r["breakdown_ms"].items() if v > 0.05})
    for s in r["secondary"]: print("   ", s["kernel"][:40], round(s["achieved"]), s["unit"], round(s["frac"], 3))
    if "cpu_baseline" in d: print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:120])
except Exception as e:
    print("ERR", e); print(open(f"gpurun_out/bench_r2b_cfg{sys.argv[1]}.err").read()[-1500:])
PY
done
