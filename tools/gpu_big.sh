set -x
mkdir -p gpurun_out
for v in 0 2; do NLB_VARIANT=$v timeout 300 python tools/bench_workload.py uniform 2097152 full_csr 5 2>&1 | tail -1; done
for v in 0 2; do NLB_VARIANT=$v timeout 600 python tools/bench_workload.py uniform 16777216 full_csr 3 2>&1 | tail -1; done
