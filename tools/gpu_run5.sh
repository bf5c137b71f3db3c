timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest11.txt 2>&1; tail -6 gpurun_out/pytest11.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; python -c "
import json; d=json.load(open('gpurun_out/bench5.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['build']['stage_ms'])"
