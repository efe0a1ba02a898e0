#!/bin/bash
# NVLink traffic of the fused export + peer-store kernel: 2 ranks on one box, rank 0 under ncu (k_export only)
O=gpurun_out
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29533 WORLD_SIZE=2
RANK=1 LOCAL_RANK=1 timeout 240 python bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > $O/r02_nvl_rank1.log 2>&1 &
RANK=0 LOCAL_RANK=0 timeout 240 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum,pcie__write_bytes.sum,pcie__read_bytes.sum \
    --clock-control none -k regex:k_export -s 6 -c 4 --csv --log-file $O/r02_nvl_export.csv \
    python bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > $O/r02_nvl_rank0.log 2>&1
wait
