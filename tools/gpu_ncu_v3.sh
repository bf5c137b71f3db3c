set -x
mkdir -p gpurun_out
python tools/profile_one.py 3 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'rowmask_kernel|emit3_kernel' -s 2 -c 2 -f -o gpurun_out/r2_v3a python tools/profile_one.py 3 > gpurun_out/ncu_v3a.log 2>&1
tail -5 gpurun_out/ncu_v3a.log
ls -la gpurun_out/*.ncu-rep
