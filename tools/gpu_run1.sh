set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest6.txt 2>&1; tail -15 gpurun_out/pytest6.txt
python tools/bench_workload.py fcc 50 full_csr 7 > gpurun_out/wl_fcc50.txt 2>&1; tail -3 gpurun_out/wl_fcc50.txt
python tools/bench_workload.py fcc 50 half_csr 7 > gpurun_out/wl_fcc50_half.txt 2>&1; tail -3 gpurun_out/wl_fcc50_half.txt
python tools/bench_workload.py uniform 2097152 full_csr 5 > gpurun_out/wl_uni2m.txt 2>&1; tail -3 gpurun_out/wl_uni2m.txt
python tools/bench_workload.py uniform 16777216 full_csr 3 > gpurun_out/wl_uni16m.txt 2>&1; tail -3 gpurun_out/wl_uni16m.txt
