set -x
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench1.json 2> gpurun_out/r2c_bench1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2c_bench_under_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
