#!/bin/bash
# One multi-GPU gpurun call: tests that need >= 2 GPUs, then bench.py both ways (torchrun ranks and
# ONE process driving all GPUs).   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_multi.sh 2'
P=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/gpu_info_g$P.csv 2>&1
echo "== pytest multi" | tee $OUT/multi_summary.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py tests/test_gpu_cpp_mirror.py -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/multi_summary.log
tail -8 $OUT/pytest_multi.log | tee -a $OUT/multi_summary.log
echo "== bench torchrun x$P" | tee -a $OUT/multi_summary.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $P > $OUT/bench_g$P.json 2> $OUT/bench_g$P.err; echo "rc=$?" | tee -a $OUT/multi_summary.log
tail -c 400 $OUT/bench_g$P.err | tee -a $OUT/multi_summary.log
echo "== bench single process x$P" | tee -a $OUT/multi_summary.log
timeout 900 python bench.py --gpus $P > $OUT/bench_sp_g$P.json 2> $OUT/bench_sp_g$P.err; echo "rc=$?" | tee -a $OUT/multi_summary.log
tail -c 400 $OUT/bench_sp_g$P.err | tee -a $OUT/multi_summary.log
if [ "$P" = "8" ]; then
  echo "== BASELINE config 4 (~100k nodes, 160 GB of matrices) x8, both process models" | tee -a $OUT/multi_summary.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --config 4 --steps 3 --warmup 2 > $OUT/bench_cfg4_g8.json 2> $OUT/bench_cfg4_g8.err; echo "rc=$?" | tee -a $OUT/multi_summary.log
  timeout 900 python bench.py --gpus 8 --config 4 --steps 3 --warmup 2 > $OUT/bench_sp_cfg4_g8.json 2> $OUT/bench_sp_cfg4_g8.err; echo "rc=$?" | tee -a $OUT/multi_summary.log
  echo "== BASELINE config 5 (IDA pattern) at 40k nodes x8, one process" | tee -a $OUT/multi_summary.log
  timeout 900 python bench.py --gpus 8 --config 5 --nodes 40000 --steps 3 --warmup 1 --jv-batched > $OUT/bench_sp_cfg5_40k_g8.json 2> $OUT/bench_sp_cfg5_40k_g8.err; echo "rc=$?" | tee -a $OUT/multi_summary.log
fi
echo "== reference arm under torchrun" | tee -a $OUT/multi_summary.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $P --steps 2 --warmup 1 > $OUT/bench_ref_g$P.json 2> $OUT/bench_ref_g$P.err; echo "rc=$?" | tee -a $OUT/multi_summary.log
echo "== done" | tee -a $OUT/multi_summary.log
