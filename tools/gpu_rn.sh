set -x
mkdir -p gpurun_out
timeout 600 python tools/v3_check.py > gpurun_out/rn_check.txt 2>&1; tail -30 gpurun_out/rn_check.txt
for v in 0 2; do NLB_VARIANT=$v timeout 120 python tools/bench_workload.py fcc 50 full_csr 9 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant',$v, d['ms_per_build'], d['stage_ms'])"; done 2>&1 | tee gpurun_out/rn_bench.txt
