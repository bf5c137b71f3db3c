# usage: LIBS="libnlist_a.so ..." bash tools/gpu_c2.sh   -> default system + 2^24 uniform stage times per library build
for lib in $LIBS; do
  export NLB200_LIB=$PWD/md_neighbor_list_b200/lib/$lib
  timeout 120 python tools/bench_workload.py fcc 50 full_csr 9 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', 'fcc50', round(d['ms_per_build']*1e3,1), {k:round(x*1e3,1) for k,x in d['stage_ms'].items()})"
  timeout 300 python tools/bench_workload.py uniform 16777216 full_csr 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', 'u16m', round(d['ms_per_build'],2), {k:round(x,2) for k,x in d['stage_ms'].items()})"
done
