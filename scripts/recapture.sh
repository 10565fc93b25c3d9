#!/bin/bash
# ncu captures only (launch list, full captures of the two hot kernels): refreshes profiles/*dram_bytes*.json
# after a source edit without the tests / sanitizers of gpu_check.sh.  gpurun -- 'bash scripts/recapture.sh'
OUT=gpurun_out; mkdir -p $OUT; export PYTHONUNBUFFERED=1
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_rows -s 2 -c 1 -f -o $OUT/prof_assemble \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_asm.log 2>&1; echo "ncu asm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bem_gemv -s 4 -c 2 -f -o $OUT/prof_gemv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_gemv.log 2>&1; echo "ncu gemv rc=$?"
