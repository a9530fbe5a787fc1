VATLQ_PASS_PHASES=1 python tools/round_cost.py 21250 125000 2>&1 | grep -v "^$" | tail -12
