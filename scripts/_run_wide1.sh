set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "smem_stage" 2>&1 | tail -5
timeout 900 python scripts/sweep_pleiades.py "W:-DB200_WIDE_WINDOW=3" "W:-DB200_WIDE_WINDOW=4" "W:-DB200_WIDE_WINDOW=5" 2>&1 | tail -8
