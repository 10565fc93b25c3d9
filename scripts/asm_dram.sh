#!/bin/bash
# Development helper: DRAM bytes of one assembly for several group sizes (ncu, two metrics only).
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for g in ${GROUPS_TO_TRY:-1 2 3 4}; do
  WBEM_ASM_GROUP=$g timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
     -k regex:k_assemble_rows -s 2 -c 1 --csv --log-file $OUT/dram_g$g.csv python scripts/quick_asm.py ${1:-20000} > /dev/null 2>&1
  echo "G=$g $(grep -E 'dram__bytes|gpu__time' $OUT/dram_g$g.csv | awk -F'\",\"' '{printf "%s=%s%s  ", $(NF-2), $NF, $(NF-1)}' | tr -d '"')"
done
