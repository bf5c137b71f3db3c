set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "window or owned_subset or absent_ghost or variants or periodic or deterministic" > gpurun_out/r2_pytest_win.txt 2>&1; tail -15 gpurun_out/r2_pytest_win.txt
