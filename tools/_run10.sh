mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pairs_plan_kernel|pass_kernel_tma' --launch-skip 700 -c 8 -o gpurun_out/u10_round21k python tools/round_cost.py 21250 > gpurun_out/u10_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_raw_summary.py gpurun_out/u10_round21k.ncu-rep > gpurun_out/u10_round21k.txt 2>&1
cut -c1-200 gpurun_out/u10_round21k.txt
