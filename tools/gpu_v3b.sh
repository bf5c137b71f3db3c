set -x
mkdir -p gpurun_out
timeout 300 python tools/v3_check.py > gpurun_out/v3_check.txt 2>&1; echo rc=$?; grep -c OK gpurun_out/v3_check.txt; grep "FAIL\|rror\|done" gpurun_out/v3_check.txt
bash tools/gpu_var.sh 0 2
for v in 0; do NLB_VARIANT=$v timeout 300 python tools/bench_workload.py uniform 2097152 full_csr 5 2>&1 | tail -1; done
for v in 0; do NLB_VARIANT=$v timeout 600 python tools/bench_workload.py uniform 16777216 full_csr 3 2>&1 | tail -1; done
