set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/halo_breakdown.py > gpurun_out/r2_halo_breakdown_n$N.txt 2>&1; grep "^{" gpurun_out/r2_halo_breakdown_n$N.txt || tail -20 gpurun_out/r2_halo_breakdown_n$N.txt
for g in 0 1; do
NLB_HALO_GRAPH=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/hg$g.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('HALO_GRAPH=$g', d['ms_per_step'], d['build'].get('ms_hot_l2_back_to_back'), d['e2e']['ms_per_step'])
"
tail -3 gpurun_out/hg$g.err
done
