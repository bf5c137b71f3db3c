set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.txt 2>&1; tail -15 gpurun_out/r2_pytest.txt
bash tools/gpu_var.sh 0 6
