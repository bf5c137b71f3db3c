timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest10.txt 2>&1; tail -6 gpurun_out/pytest10.txt
for v in 2 3; do echo "== variant $v"; NLB_VARIANT=$v timeout 300 python tools/bench_workload.py fcc 50 full_csr 9 2>&1 | tail -1 | cut -c1-700; done
echo "== half"; timeout 300 python tools/bench_workload.py fcc 50 half_csr 9 2>&1 | tail -1 | cut -c1-700
for v in 2 3; do echo "== 2M variant $v"; NLB_VARIANT=$v timeout 300 python tools/bench_workload.py uniform 2097152 full_csr 5 2>&1 | tail -1 | cut -c1-700; done
