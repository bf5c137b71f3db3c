# A/B of the emission kernels on the default system (variant 0 = emitwin, 9 = emitrun) + the GPU tests
set -x
mkdir -p gpurun_out
bash tools/gpu_var.sh 0 9 0 9 2>&1 | tee gpurun_out/ew_ab.txt
DENSITY=0.5 bash tools/gpu_var.sh 0 9 2>&1 | tee -a gpurun_out/ew_ab.txt
for v in 0 9; do NLB_VARIANT=$v timeout 300 python tools/bench_workload.py uniform 2097152 full_csr 5 2>&1 | tail -1 | cut -c1-900 | tee -a gpurun_out/ew_ab.txt; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/ew_pytest.txt
