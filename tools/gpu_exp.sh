# one experiment round on the GPU box: parity smoke, then the default-system bench per library variant, kernel variant
# and PDL setting
set -u
timeout 150 bash tools/gpu_quick.sh 2>&1 | tail -4
cp md_neighbor_list_b200/lib/libnlist_b200.so /tmp/base.so
for v in base $(ls md_neighbor_list_b200/lib/variants/ 2>/dev/null); do
  if [ "$v" != "base" ]; then cp md_neighbor_list_b200/lib/variants/$v md_neighbor_list_b200/lib/libnlist_b200.so; fi
  for pdl in ${PDLS:-0}; do
    for var in ${VARIANTS:-0}; do
      for mode in ${MODES:-full_csr}; do
        echo "== $v pdl=$pdl variant=$var $mode"
        NLB_VARIANT=$var NLB200_PDL=$pdl timeout 300 python tools/bench_workload.py ${WORKLOAD:-fcc 50} $mode ${REPS:-15} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_build'], d['entries'], d['stage_ms'])"
      done
    done
  done
done
cp /tmp/base.so md_neighbor_list_b200/lib/libnlist_b200.so
