# round-2 validation + evidence on one B200: smoke, GPU tests, C++ driver self-tests, both bench arms, then (each only
# after its plain command exited 0) the ncu launch list of the bench command, one ncu --set full capture of every kernel
# of a default-system build, and the per-kernel DRAM traffic of a build of 2^24 uniform particles
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r02_pytest_gpu.txt
./drivers/make_list_b200.out gpu 1.0 100 1 > gpurun_out/driver_gpu.txt 2>&1; tail -2 gpurun_out/driver_gpu.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/r02_bench_ref.json
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/bench_final.err; rc=$?; tail -c 600 gpurun_out/r02_bench_final.json; tail -2 gpurun_out/bench_final.err
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
fi
timeout 300 python tools/profile_one.py 3 > gpurun_out/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -s 14 -c 7 -f -o gpurun_out/r02_build_full python tools/profile_one.py 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 300 python tools/profile_uniform16m.py > gpurun_out/plain_u16.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_u16m_traffic.csv python tools/profile_uniform16m.py > gpurun_out/ncu_u16.log 2>&1
tail -2 gpurun_out/ncu_u16.log
timeout 300 python tools/profile_uniform16m.py > /dev/null 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"emitwin_kernel" -s 1 -c 1 -f -o gpurun_out/r02_u16m_emitwin python tools/profile_uniform16m.py > gpurun_out/ncu_u16_ew.log 2>&1
tail -2 gpurun_out/ncu_u16_ew.log
