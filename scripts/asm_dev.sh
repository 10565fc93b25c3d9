#!/bin/bash
# Development helper (one gpurun call): the stream kernel against the oracle and the colour-per-launch kernel
# at two sizes, under a timeout (a dependency bug would otherwise spin), the group-size sweep, the build variants.
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== $(date)" > $OUT/asm_dev.log
timeout 120 python scripts/quick_asm.py 1500 --check 2>&1 | tail -2 | tee -a $OUT/asm_dev.log
timeout 300 python scripts/quick_asm.py ${1:-20000} --check 2>&1 | tail -2 | tee -a $OUT/asm_dev.log
timeout 120 python scripts/quick_asm.py ${1:-20000} --variant 2 2>&1 | tail -1 | tee -a $OUT/asm_dev.log
for g in 1 2 3 4 6; do
  echo "WBEM_ASM_GROUP=$g" | tee -a $OUT/asm_dev.log
  WBEM_ASM_GROUP=$g timeout 120 python scripts/quick_asm.py ${1:-20000} 2>&1 | tail -1 | tee -a $OUT/asm_dev.log
done
for lib in wavebem_b200/lib/libwbem_*.so; do
  [ -f $lib ] || continue
  case $lib in *probes*|*waitlog*) continue;; esac
  for g in ${VARIANT_GROUPS:-3}; do
    echo "$lib G=$g" | tee -a $OUT/asm_dev.log
    WBEM_ASM_GROUP=$g WBEM_LIB=$PWD/$lib timeout 300 python scripts/quick_asm.py ${1:-20000} --check 2>&1 | tail -2 | tee -a $OUT/asm_dev.log
  done
done
if [ -f wavebem_b200/lib/libwbem_waitlog.so ]; then
  for g in 1 2 3; do echo "waitlog G=$g" | tee -a $OUT/asm_dev.log
    WBEM_ASM_GROUP=$g WBEM_LIB=$PWD/wavebem_b200/lib/libwbem_waitlog.so timeout 120 python scripts/asm_waitlog.py ${1:-20000} 2>&1 | head -6 | tee -a $OUT/asm_dev.log
  done
fi
