#!/bin/bash
# weak + strong scaling at N = 2, 4, 8 on one box (N = 1 is the default bench run of tools/r02_profiles.sh)
O=gpurun_out
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
      bench.py --gpus $N --steps 6 --warmup 3 > $O/r02_scale_$N.json 2> $O/r02_scale_$N.err
done
