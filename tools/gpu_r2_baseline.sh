# round-2 step 1: the new parity tests on the round-1 kernels, the reference GPU path's self-test + timing on this box,
# and a bench line (before any kernel changes)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_base.txt 2>&1; tail -5 gpurun_out/r2_pytest_base.txt
for v in warp_unroll_smem warp_unroll smem_mesh ref; do
  ( time timeout 300 ./oracle/_ref/gpu/make_list_gpu_$v.out 128 7 ) > gpurun_out/r2_refgpu_$v.txt 2>&1
  tail -6 gpurun_out/r2_refgpu_$v.txt
done
python bench.py --steps 30 --warmup 5 > gpurun_out/r2_bench_base.json 2> gpurun_out/r2_bench_base.err; tail -c 2500 gpurun_out/r2_bench_base.json; tail -3 gpurun_out/r2_bench_base.err
