set -x
python -m pytest tests/test_gpu_parity.py -x -q -k "staged_saveat or smem_stage" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench1.json 2> gpurun_out/r2c_bench1.err
tail -c 400 gpurun_out/r2c_bench1.json | head -c 300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2c_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k b200_integrate -s 1 -c 1 -f -o gpurun_out/prof_r2c_pleiades python scripts/prof_pleiades.py "-DB200_WIDE=1" > gpurun_out/prof_r2c_p.log 2>&1
ls -la gpurun_out/prof_r2c_* gpurun_out/r2c_*
