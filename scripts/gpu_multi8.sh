#!/bin/bash
# 8-GPU run of the scaling point and BASELINE config 4: gpurun --gpus 8 -- 'bash scripts/gpu_multi8.sh [all]'
set -u
MODE=${1:-short}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== bench --gpus 8 (weak scaling point, N ~ 57k)" | tee $OUT/multi8_summary.log
NCCL_DEBUG=WARN timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 > $OUT/bench_g8.json 2> $OUT/bench_g8.err
echo "rc=$?" | tee -a $OUT/multi8_summary.log; tail -1 $OUT/bench_g8.json | cut -c1-600 | tee -a $OUT/multi8_summary.log
echo "== bench --gpus 4 (N ~ 40k, config 3 size)" | tee -a $OUT/multi8_summary.log
NCCL_DEBUG=WARN timeout 600 $TR --nproc-per-node 4 --master-port 29525 bench.py --gpus 4 > $OUT/bench_g4.json 2> $OUT/bench_g4.err
echo "rc=$?" | tee -a $OUT/multi8_summary.log; tail -1 $OUT/bench_g4.json | cut -c1-600 | tee -a $OUT/multi8_summary.log
echo "== bench --gpus 8 --nodes 35355 (config 4: N ~ 100k, 160 GB of matrices)" | tee -a $OUT/multi8_summary.log
NCCL_DEBUG=WARN timeout 900 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --nodes 35355 --steps 3 --warmup 3 > $OUT/bench_g8_100k.json 2> $OUT/bench_g8_100k.err
echo "rc=$?" | tee -a $OUT/multi8_summary.log; tail -1 $OUT/bench_g8_100k.json | cut -c1-600 | tee -a $OUT/multi8_summary.log
if [ "$MODE" = "all" ]; then
  echo "== bench --gpus 8 --nodes 14142 (config 3: N ~ 40k on 8 GPUs)" | tee -a $OUT/multi8_summary.log
  NCCL_DEBUG=WARN timeout 600 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --nodes 14142 > $OUT/bench_g8_40k.json 2> $OUT/bench_g8_40k.err
  echo "rc=$?" | tee -a $OUT/multi8_summary.log; tail -1 $OUT/bench_g8_40k.json | cut -c1-600 | tee -a $OUT/multi8_summary.log
  echo "== IDA call pattern, N ~ 40k, 8 GPUs (config 5)" | tee -a $OUT/multi8_summary.log
  timeout 600 $TR --nproc-per-node 8 --master-port 29524 scripts/ida_emulation.py --nodes 40000 --steps 3 > $OUT/ida_g8_40k.json 2> $OUT/ida_g8_40k.err
  echo "rc=$?" | tee -a $OUT/multi8_summary.log; tail -1 $OUT/ida_g8_40k.json | tee -a $OUT/multi8_summary.log
fi
