#!/bin/bash
# Development: A/B builds of ONE translation unit (UNIT, default k_clqr.cu) linked against the current objects of every other
# unit, written to lq_mpc_b200/_lib/variants/<name>.so; select one at run time with LQMPC_LIB=<path>.
# usage: scripts/unit_variants.sh name1 "-DLQ_K2_PF_SMALL=0 -DLQ_K2_MINB_SMALL=4" name2 "..." ...
set -e
cd "$(dirname "$0")/.."
python -c "import lq_mpc_b200._build as b; b.build()"
OBJ=lq_mpc_b200/_lib/obj
OUT=lq_mpc_b200/_lib/variants
mkdir -p $OUT
UNIT=${UNIT:-k_clqr.cu}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr"
while [ $# -ge 2 ]; do
  name=$1; extra=$2; shift 2
  (
    nvcc $FLAGS $extra -c lq_mpc_b200/csrc/$UNIT -o $OUT/$name.unit.o 2> $OUT/$name.ptxas.log
    others=$(ls $OBJ/*.o | grep -v $UNIT.o)
    nvcc -shared -o $OUT/$name.so $OUT/$name.unit.o $others -cudart static -gencode arch=compute_100a,code=sm_100a
    rm -f $OUT/$name.unit.o
    grep -A2 "mpc_solve_kernelILi2ELi1ELb0\|simulate_kernelILi2ELi1ELb0\|bounds_kernelILi2ELi1E" $OUT/$name.ptxas.log | grep Used | sed "s/^/$name: /"
  ) &
done
wait
