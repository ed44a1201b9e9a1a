mkdir -p gpurun_out
# launch list of a short default bench run (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 4000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r2_bench.log 2>&1
echo "launch list exit=$?"
# full captures: one layer's GEMMs + attention, then sampler / remask / embed gather, then the codec stack
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|attention_tcgen05" -s 52 -c 6 -f -o gpurun_out/prof_layer_r2 python tools/profile_step.py > gpurun_out/ncu_layer_r2.log 2>&1
echo "ncu layer exit=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sample_rows|remask|embed_gather" -s 3 -c 4 -f -o gpurun_out/prof_misc_r2 python tools/profile_step.py > gpurun_out/ncu_misc_r2.log 2>&1
echo "ncu misc exit=$?"
timeout 400 ncu --clock-control none -k regex:"conv_tcgen05|rvq|codec_" -c 40 --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy -f -o gpurun_out/prof_conv_r2 python tools/profile_step.py --codec > gpurun_out/ncu_conv_r2.log 2>&1
echo "ncu conv exit=$?"
timeout 300 python -m pytest tests/test_gpu_attention.py -q -s 2>&1 | tail -5
