python -m pytest tests -m gpu -x -q > gpurun_out/pytest7.txt 2>&1; tail -5 gpurun_out/pytest7.txt
python tools/bench_workload.py fcc 50 full_csr 7 > gpurun_out/wl_fcc50.txt 2>&1; tail -1 gpurun_out/wl_fcc50.txt
python tools/bench_workload.py fcc 50 half_csr 7 > gpurun_out/wl_fcc50_half.txt 2>&1; tail -1 gpurun_out/wl_fcc50_half.txt
python tools/bench_workload.py uniform 2097152 full_csr 5 > gpurun_out/wl_uni2m.txt 2>&1; tail -1 gpurun_out/wl_uni2m.txt
