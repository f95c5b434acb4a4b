set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err
tail -c 600 gpurun_out/r2b_bench1.json | head -c 300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2b_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k b200_integrate -s 2 -c 1 -f -o gpurun_out/prof_r2d_saveat python scripts/prof_dev.py f64 saveat > gpurun_out/prof_r2d_1.log 2>&1
ncu --set full --clock-control none --import-source on -k b200_integrate -s 2 -c 1 -f -o gpurun_out/prof_r2d_final python scripts/prof_dev.py f64 > gpurun_out/prof_r2d_2.log 2>&1
ls -la gpurun_out/prof_r2d_*
