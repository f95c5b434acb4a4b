set -x
for w in rodas5p ros23; do
  ncu --set full --clock-control none --import-source on -k b200_integrate -s 1 -c 1 -f -o gpurun_out/prof_r2c_$w python scripts/prof_generic.py $w 262144 > gpurun_out/prof_r2c_$w.log 2>&1
done
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_r2c.py > gpurun_out/sanitize_r2c_$tool.log 2>&1
  tail -3 gpurun_out/sanitize_r2c_$tool.log
done
