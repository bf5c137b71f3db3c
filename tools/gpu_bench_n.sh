# usage: bash tools/gpu_bench_n.sh N
set -x
N=$1
mkdir -p gpurun_out
( time python -X faulthandler -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -c 3500 gpurun_out/r2_bench_n$N.json; tail -8 gpurun_out/r2_bench_n$N.err
