timeout 600 python -m pytest tests -m gpu -x -q -k "clustered or cell_capacity or default_system_full" 2>&1 | tail -3
timeout 900 python tools/sweep_c4.py 18 gpurun_out/sweep_c4_n18.jsonl 2>&1 | tail -14
