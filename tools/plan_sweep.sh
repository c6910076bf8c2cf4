#!/bin/bash
# pass-1 work plan variants at the shard shapes (tools/shard_shape.py): unit of the cuts (stage | sub-chunk) x cost of the last tile
mkdir -p gpurun_out
for cfg in "stage 1.0" "subchunk 1.0" "subchunk 0.45" "subchunk 0.6" "subchunk 0.8" "stage 0.45"; do
    set -- $cfg
    echo "== unit $1, last-tile cost $2"
    CPD_B200_PLAN_UNIT=$1 CPD_B200_PLAN_LAST_COST=$2 timeout 300 python tools/shard_shape.py 2>&1 | grep -E "shard 1/(4|8)"
done
