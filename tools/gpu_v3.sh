set -x
mkdir -p gpurun_out
timeout 300 python tools/v3_check.py > gpurun_out/v3_check.txt 2>&1; echo rc=$?; grep -c OK gpurun_out/v3_check.txt; grep "FAIL\|rror\|done" gpurun_out/v3_check.txt
bash tools/gpu_var.sh 0 5 2
