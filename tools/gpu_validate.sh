set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.txt 2>&1; tail -5 gpurun_out/r2_pytest.txt
( time python bench.py --steps 30 --warmup 5 ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 1500 gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err
