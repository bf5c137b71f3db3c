# run-mask unit size variants on the default system; emission kernels (9 = gather, 10 = window) on large systems
set -x
mkdir -p gpurun_out
VARIANTS="0" bash tools/gpu_variants.sh 2>&1 | grep -v "^+" | tee gpurun_out/exp2_rn.txt
for v in 9 10; do
  echo "== uniform 2^24 variant $v"; NLB_VARIANT=$v timeout 300 python tools/bench_workload.py uniform 16777216 full_csr 3 2>&1 | tail -1 | cut -c1-1000
  echo "== fcc L=160 variant $v"; NLB_VARIANT=$v timeout 300 python tools/bench_workload.py fcc 160 full_csr 3 2>&1 | tail -1 | cut -c1-1000
  echo "== uniform 2^19 variant $v"; NLB_VARIANT=$v timeout 300 python tools/bench_workload.py uniform 524288 full_csr 5 2>&1 | tail -1 | cut -c1-1000
done 2>&1 | tee gpurun_out/exp2_emit.txt
