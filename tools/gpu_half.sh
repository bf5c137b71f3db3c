timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 0 4; do echo "== half variant $v"; NLB_VARIANT=$v timeout 300 python tools/bench_workload.py fcc 50 half_csr 9 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_build'], d['stage_ms'])"; done
echo "== full"; timeout 300 python tools/bench_workload.py fcc 50 full_csr 9 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_build'], d['stage_ms'])"
./drivers/make_list_b200.out cpu 1.0 20 1 2>&1 | tail -2
