#!/bin/bash
# Development helper: ncu --set full of the single-launch assembly kernel (one gpurun call).
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_rows -s 2 -c 1 -f -o $OUT/prof_stream \
   python scripts/quick_asm.py ${1:-20000} > $OUT/ncu_stream.log 2>&1; echo "ncu rc=$?"
tail -3 $OUT/ncu_stream.log
