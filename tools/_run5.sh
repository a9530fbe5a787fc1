mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_coreset.py tests/test_gpu_prune.py tests/test_gpu_scale.py tests/test_gpu_query.py -x -q > gpurun_out/u5_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/u5_tests.log
python tools/round_cost.py 21250 125000 170000 > gpurun_out/u5_round_cost.json 2>/dev/null; echo "rc=$?"
VATLQ_PLAN_SMEM=0 python tools/round_cost.py 21250 125000 > gpurun_out/u5_round_cost_l2plan.json 2>/dev/null; echo "rc=$?"
echo SMEM; cat gpurun_out/u5_round_cost.json; echo L2; cat gpurun_out/u5_round_cost_l2plan.json
