set -x
mkdir -p gpurun_out
( time python bench.py --steps 30 --warmup 5 ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 6000 gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err
