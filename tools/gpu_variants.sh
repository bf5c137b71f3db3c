# runs the default-system bench with each library variant in md_neighbor_list_b200/lib/variants
cp md_neighbor_list_b200/lib/libnlist_b200.so /tmp/base.so
for v in base $(ls md_neighbor_list_b200/lib/variants/); do
  if [ "$v" != "base" ]; then cp md_neighbor_list_b200/lib/variants/$v md_neighbor_list_b200/lib/libnlist_b200.so; fi
  for var in ${VARIANTS:-0}; do
    echo "== $v variant=$var"
    NLB_VARIANT=$var python tools/bench_workload.py fcc 50 full_csr 9 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_build'], d['stage_ms'])"
  done
done
cp /tmp/base.so md_neighbor_list_b200/lib/libnlist_b200.so
