mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/final_pytest.txt
for c in 2 1 3 4; do timeout 600 python bench.py --config $c > gpurun_out/final_bench_cfg$c.json 2> gpurun_out/final_bench_cfg$c.err; done
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_ref.json 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1
tail -3 gpurun_out/final_pytest.txt; cat gpurun_out/final_smoke.txt | tail -2
