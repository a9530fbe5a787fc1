mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_next.py tests/test_gpu_query.py -x -q > gpurun_out/u9_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/u9_tests.log
timeout 300 python tools/time_next_rows.py > gpurun_out/u9_next_rows.json 2>/dev/null; echo "rc=$?"; cat gpurun_out/u9_next_rows.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 40 --csv --log-file gpurun_out/u9_launches_21k.csv python tools/round_cost.py 21250 > /dev/null 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/u9_launches_21k.csv | cut -d, -f5,15 | head -50
