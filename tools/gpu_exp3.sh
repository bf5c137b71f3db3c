# emitwin expansion loop: 2 vs 4 entries per trip (variant 10 = emitwin forced)
mkdir -p gpurun_out
cp md_neighbor_list_b200/lib/libnlist_b200.so /tmp/base.so
for v in base $(ls md_neighbor_list_b200/lib/variants/); do
  if [ "$v" != "base" ]; then cp md_neighbor_list_b200/lib/variants/$v md_neighbor_list_b200/lib/libnlist_b200.so; fi
  echo "== $v default system, emitwin forced"; NLB_VARIANT=10 python tools/bench_workload.py fcc 50 full_csr 9 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_build'], d['stage_ms']['emit_run'])"
  echo "== $v uniform 2^21"; python tools/bench_workload.py uniform 2097152 full_csr 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_build'], d['stage_ms']['emit_run'])"
  echo "== $v uniform 2^24"; python tools/bench_workload.py uniform 16777216 full_csr 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_build'], d['stage_ms']['emit_run'])"
done 2>&1 | tee gpurun_out/exp3.txt
cp /tmp/base.so md_neighbor_list_b200/lib/libnlist_b200.so
timeout 120 ./drivers/make_list_b200.out slab 1 1.0 5; timeout 200 ./drivers/make_list_b200.out slab 2 1.0 5
