# usage: bash tools/mkvariant.sh NAME "-DFLAG=.. ..."   -> md_neighbor_list_b200/lib/libnlist_NAME.so (tuning builds)
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DNLB_WAIT_GUARD -Xcompiler -fPIC -Xcompiler -O3 $2 \
  -shared -o md_neighbor_list_b200/lib/libnlist_$1.so md_neighbor_list_b200/csrc/nlist_api.cu md_neighbor_list_b200/csrc/workloads.cpp
