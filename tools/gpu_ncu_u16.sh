# usage: bash tools/gpu_ncu_u16.sh <kernel-regex> <out-name> [skip] [count]   (2^24 uniform particles)
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-1} -c ${4:-1} -f -o gpurun_out/$2 python tools/profile_uniform16m.py > gpurun_out/ncu_$2.log 2>&1
tail -3 gpurun_out/ncu_$2.log
