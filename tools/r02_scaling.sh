#!/bin/bash
# weak + strong scaling at N = 2, 4, 8 on one box (N = 1 is the default bench run); NVLink counters around the N = 8 run
O=gpurun_out
for N in 2 4 8; do
  if [ $N = 8 ]; then nvidia-smi nvlink -gt d -i 0 > $O/r02_nvlink_before.txt 2>&1; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
      bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_scale_$N.json 2> $O/r02_scale_$N.err
  if [ $N = 8 ]; then nvidia-smi nvlink -gt d -i 0 > $O/r02_nvlink_after.txt 2>&1; fi
done
