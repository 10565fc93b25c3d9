#!/bin/bash
# Multi-GPU check: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'
set -u
NG=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > $OUT/multi_gpu_info.csv 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/multi_build.log 2>&1
echo "== pytest multi" | tee $OUT/multi_summary.log
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/multi_summary.log
tail -15 $OUT/pytest_multi.log | tee -a $OUT/multi_summary.log
if [ "$NG" != "1" ]; then
  echo "== bench --gpus $NG --no-p2p (ncclAllGather)" | tee -a $OUT/multi_summary.log
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29516 \
      bench.py --gpus $NG --no-p2p > $OUT/bench_g${NG}_nccl.json 2> $OUT/bench_g${NG}_nccl.err
  echo "rc=$?" | tee -a $OUT/multi_summary.log
  tail -1 $OUT/bench_g${NG}_nccl.json | cut -c1-1500 | tee -a $OUT/multi_summary.log
fi
for n in 1 $NG; do
  echo "== bench --gpus $n" | tee -a $OUT/multi_summary.log
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --no-cpu-baseline > $OUT/bench_g$n.json 2> $OUT/bench_g$n.err
  else
    NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n > $OUT/bench_g$n.json 2> $OUT/bench_g$n.err
  fi
  echo "rc=$?" | tee -a $OUT/multi_summary.log
  tail -1 $OUT/bench_g$n.json | cut -c1-1500 | tee -a $OUT/multi_summary.log
  tail -3 $OUT/bench_g$n.err | tee -a $OUT/multi_summary.log
done
