#!/bin/bash
# Development helper: time the assembly kernel variants built by build.build_variant()
# (one gpurun call):   gpurun -- 'bash scripts/asm_variants.sh 20000'
N=${1:-20000}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== $(date)" >> $OUT/asm_variants.log
for lib in wavebem_b200/lib/libwbem.so wavebem_b200/lib/libwbem_*.so; do
  [ -f $lib ] || continue
  WBEM_LIB=$PWD/$lib timeout 300 python scripts/quick_asm.py $N --check 2>&1 | tail -2 | tee -a $OUT/asm_variants.log
done
