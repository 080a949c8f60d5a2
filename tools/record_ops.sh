#!/bin/bash
# ncu --set full of the config-4 operator kernels (Voxelization / DynamicScatter / scatter_v2 and the CUB sort they use)
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
ncu --set full --clock-control none -k regex:"k_|DeviceRadixSort|DeviceScan" -s 60 -c 24 -f -o $out/${tag}_ops \
    python tools/bench_ops.py > $out/${tag}_ncu_ops.log 2>&1
tail -3 $out/${tag}_ncu_ops.log | cut -c1-200
