set -x
mkdir -p gpurun_out
for v in 0 2; do NLB_VARIANT=$v timeout 600 python tools/scale_bench.py --steps 3 2>&1 | tail -1 | cut -c1-900; done
for v in 0 2; do
  NLB_VARIANT=$v timeout 300 python tools/profile_uniform16m.py > gpurun_out/plain_u16_$v.log 2>&1 &&
  NLB_VARIANT=$v ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_u16m_traffic_v$v.csv python tools/profile_uniform16m.py > gpurun_out/ncu_u16_$v.log 2>&1
  tail -2 gpurun_out/ncu_u16_$v.log
done
