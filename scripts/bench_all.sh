#!/bin/bash
# All BASELINE.json configurations on one GPU (value only for the secondary ones).
set -x
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err
for w in lorenz_tsit5_saveat_1m_f32 lorenz_tsit5_final_1m robertson_rodas5p_1m robertson_rosenbrock23_1m pleiades_vern7_256k; do
  python bench.py --workload $w --steps 5 --warmup 2 --cpu-seconds 6 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
python bench.py --workload lorenz_sweep_64m --steps 3 --warmup 1 > gpurun_out/bench_sweep64m_1gpu.json 2> gpurun_out/bench_sweep.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
