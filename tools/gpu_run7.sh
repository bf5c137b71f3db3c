bash tools/gpu_quick.sh 2>&1 | tail -3 | cut -c1-200
for n in 2097152 16777216; do timeout 400 python tools/bench_workload.py uniform $n full_csr 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n'], d['ms_per_build'], d['stage_ms'])"; done
