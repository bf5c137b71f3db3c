set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/halo_breakdown.py > gpurun_out/r2_halo_p2p_n$N.txt 2>&1; grep "^{" gpurun_out/r2_halo_p2p_n$N.txt || tail -20 gpurun_out/r2_halo_p2p_n$N.txt
NLB_HALO=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/halo_breakdown.py 2>&1 | grep "^{"
timeout 600 bash tools/gpu_bench_n.sh $N
