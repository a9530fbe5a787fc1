mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_coreset.py tests/test_gpu_prune.py tests/test_gpu_scale.py tests/test_gpu_kmeans.py -x -q > gpurun_out/u12_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/u12_tests.log
VATLQ_PASS_PHASES=1 python tools/round_cost.py 21250 125000 170000 2>&1 | grep -v "^$" | grep -v "47 rounds\|47 launches\|57 rounds\|57 launches" | tail -12
