#!/bin/bash
# Run under `gpurun --gpus 2`: does capping NCCL's CTAs (fewer SMs taken from the gridder while a collective overlaps it)
# help the pipelined continuum step?  Measured on 2 x B200 (round 1): default 2.68 ms/step, NCCL_MAX_CTAS=16 2.78,
# 8 2.99, 4 3.81, 2 7.26 -- no: the collectives' SM-time, not their width, is what the step pays for; left at the default.
for c in default 2 4 8 16; do
  if [ "$c" = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$c; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('CTAS=$c', d['ms_per_step'], d['value']/1e9)"
done
