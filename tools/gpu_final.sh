# round-end validation on the GPU box: full GPU test suite, C++ driver self-tests, both bench arms, the ncu launch list
# of the bench command and one full ncu capture of the build's kernels (each only after its command exited 0 plainly)
set -u
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.txt 2>&1; tail -3 gpurun_out/pytest_final.txt
./drivers/make_list_b200.out gpu 1.0 100 1 > gpurun_out/driver_gpu.txt 2>&1; tail -2 gpurun_out/driver_gpu.txt
./drivers/make_list_b200.out cpu 0.5 20 1 > gpurun_out/driver_cpu.txt 2>&1; tail -2 gpurun_out/driver_cpu.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; rc=$?; tail -c 1500 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
fi
timeout 300 python tools/profile_one.py 3 > gpurun_out/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pairmask_kernel|emit_kernel|rowcount_kernel' -s 6 -c 3 -o gpurun_out/prof_final2 -f python tools/profile_one.py 3 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
