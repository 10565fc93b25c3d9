#!/bin/bash
# 8-GPU run (one gpurun call, charged 8x): BASELINE config 3 (40k nodes, strong scaling point) both process
# models, config 4 (100k nodes, 160 GB of matrices), the rank-sharded tests, the reference arm under torchrun.
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_multi8.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== bench torchrun x8 (config 3, N ~ 40k)" | tee $OUT/multi8_summary.log
timeout 300 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 > $OUT/bench_g8.json 2> $OUT/bench_g8.err; echo "rc=$?" | tee -a $OUT/multi8_summary.log
echo "== bench one process x8 (config 3)" | tee -a $OUT/multi8_summary.log
timeout 300 python bench.py --gpus 8 --no-cpu-baseline > $OUT/bench_sp_g8.json 2> $OUT/bench_sp_g8.err; echo "rc=$?" | tee -a $OUT/multi8_summary.log
echo "== bench torchrun x4 (config 3)" | tee -a $OUT/multi8_summary.log
timeout 300 $TR --nproc-per-node 4 --master-port 29525 bench.py --gpus 4 --no-cpu-baseline > $OUT/bench_g4.json 2> $OUT/bench_g4.err; echo "rc=$?" | tee -a $OUT/multi8_summary.log
echo "== config 4 (N ~ 100k) torchrun x8" | tee -a $OUT/multi8_summary.log
timeout 400 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --config 4 --steps 3 --warmup 2 --no-cpu-baseline > $OUT/bench_cfg4_g8.json 2> $OUT/bench_cfg4_g8.err; echo "rc=$?" | tee -a $OUT/multi8_summary.log
echo "== rank-sharded tests (8 ranks)" | tee -a $OUT/multi8_summary.log
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu > $OUT/pytest_multi8.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/multi8_summary.log
tail -3 $OUT/pytest_multi8.log | tee -a $OUT/multi8_summary.log
echo "== reference arm under torchrun x8" | tee -a $OUT/multi8_summary.log
timeout 300 $TR --nproc-per-node 8 --master-port 29532 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $OUT/bench_ref_g8.json 2> $OUT/bench_ref_g8.err; echo "rc=$?" | tee -a $OUT/multi8_summary.log
echo "== done" | tee -a $OUT/multi8_summary.log
