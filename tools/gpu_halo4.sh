set -x
mkdir -p gpurun_out
N=${1:-2}
for f in 1 0; do
NLB_HALO_FUSED=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$f bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/hf$f.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('FUSED=$f', d['ms_per_step'], d['build'].get('ms_hot_l2_back_to_back'), d['e2e']['ms_per_step'], 'rows_ok', d.get('rank0_rows_match_oracle'), 'c3', d.get('c3',{}).get('ms_per_build'), d.get('c3',{}).get('weak_scaling_efficiency'))
"
tail -3 gpurun_out/hf$f.err
done
