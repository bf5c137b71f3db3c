# round-2 evidence: (1) per-launch durations of the bench command, (2) one ncu --set full capture of every kernel of a
# build of the default system (third build of tools/profile_one.py)
set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python tools/profile_one.py 3 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 14 -c 7 -f -o gpurun_out/r02_build_full python tools/profile_one.py 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
