# alternative builds of the library (NLB200_LIB) x kernel variants: default-system stage times
# usage: LIBS="libnlist_a.so libnlist_b.so" VARS="0" bash tools/gpu_tune.sh
for lib in $LIBS; do
  for v in ${VARS:-0}; do
    NLB200_LIB=$PWD/md_neighbor_list_b200/lib/$lib NLB_VARIANT=$v timeout 120 python tools/bench_workload.py fcc 50 full_csr 9 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', 'variant', $v, round(d['ms_per_build']*1e3,1), {k:round(x*1e3,1) for k,x in d['stage_ms'].items()})"
  done
done
