set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest9.txt 2>&1; tail -4 gpurun_out/pytest9.txt
./drivers/make_list_b200.out gpu 1.0 100 1 > gpurun_out/driver_gpu.txt 2>&1; tail -3 gpurun_out/driver_gpu.txt
./drivers/make_list_b200.out cpu 0.5 20 1 > gpurun_out/driver_cpu.txt 2>&1; tail -3 gpurun_out/driver_cpu.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -c 3000 gpurun_out/bench4.json; tail -3 gpurun_out/bench4.err
