# usage: bash tools/gpu_ncu_k.sh <kernel-regex> <out-name> [launch-skip] [count]
set -x
mkdir -p gpurun_out
python tools/profile_one.py 3 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-1} -c ${4:-1} -f -o gpurun_out/$2 python tools/profile_one.py 3 > gpurun_out/ncu_$2.log 2>&1
tail -3 gpurun_out/ncu_$2.log
