#!/bin/bash
# One gpurun call: smoke (under compute-sanitizer first), GPU tests, bench, ncu launch list + full captures.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [quick|full]'
set -u
MODE=${1:-full}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_info.csv 2>&1
echo "== build + smoke" | tee $OUT/summary.log
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.log
tail -3 $OUT/smoke.log | tee -a $OUT/summary.log
if [ "$MODE" != "quick" ]; then
  echo "== compute-sanitizer memcheck (smoke)" | tee -a $OUT/summary.log
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/summary.log
  grep -E "ERROR SUMMARY|Invalid|out of bounds" $OUT/memcheck.log | head -5 | tee -a $OUT/summary.log
  echo "== compute-sanitizer racecheck (smoke)" | tee -a $OUT/summary.log
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/summary.log
  grep -E "RACECHECK SUMMARY|hazard|ERROR" $OUT/racecheck.log | head -8 | tee -a $OUT/summary.log
fi
echo "== pytest -m gpu" | tee -a $OUT/summary.log
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.log
tail -40 $OUT/pytest_gpu.log | tee -a $OUT/summary.log
echo "== bench" | tee -a $OUT/summary.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/summary.log
cat $OUT/bench.json | tee -a $OUT/summary.log; tail -5 $OUT/bench.err | tee -a $OUT/summary.log
if [ "$MODE" != "quick" ]; then
  echo "== ncu launch list" | tee -a $OUT/summary.log
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/summary.log
  echo "== ncu full: assembly + gemv" | tee -a $OUT/summary.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_assemble_rows -s 2 -c 1 -f -o $OUT/prof_assemble \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_asm.log 2>&1; echo "ncu asm rc=$?" | tee -a $OUT/summary.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bem_gemv -s 4 -c 2 -f -o $OUT/prof_gemv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_gemv.log 2>&1; echo "ncu gemv rc=$?" | tee -a $OUT/summary.log
fi
if [ "$MODE" != "quick" ]; then
  echo "== BASELINE configs 3 (one GPU) and 5 (IDA pattern, single and batched J.v)" | tee -a $OUT/summary.log
  timeout 600 python bench.py --config 3 --steps 3 --no-cpu-baseline 2>/dev/null | tail -1 > $OUT/bench_cfg3_g1.json; echo "cfg3 rc=$?" | tee -a $OUT/summary.log
  timeout 600 python bench.py --config 5 --nodes 20000 --steps 3 --warmup 1 2>/dev/null | tail -1 > $OUT/bench_cfg5_20k.json
  timeout 600 python bench.py --config 5 --nodes 20000 --steps 3 --warmup 1 --jv-batched 2>/dev/null | tail -1 > $OUT/bench_cfg5_20k_batched.json
  timeout 600 python bench.py --config 5 --steps 10 --warmup 2 2>/dev/null | tail -1 > $OUT/bench_cfg5_4k.json
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > $OUT/bench_reference.json
  echo "== DRAM bytes of one assembly against the row-tile group size G (WBEM_ASM_GROUP)" | tee -a $OUT/summary.log
  GROUPS_TO_TRY="1 2 3 4" bash scripts/asm_dram.sh 20000 2>&1 | tee $OUT/assemble_dram_vs_group.txt | tee -a $OUT/summary.log
fi
echo "== done" | tee -a $OUT/summary.log
