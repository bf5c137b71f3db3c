# usage: bash tools/gpu_halo5.sh N ["1 0"]   contract bench at N GPUs without the extras; folded (1) / separate (0) packing
set -x
mkdir -p gpurun_out
N=${1:-4}
for f in ${2:-1 0}; do
NLB_HALO_FUSED=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$f bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>gpurun_out/hf$f.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=$N FUSED=$f', d['ms_per_step'], d['build'].get('ms_hot_l2_back_to_back'), d['e2e']['ms_per_step'])
"
tail -3 gpurun_out/hf$f.err
done
