#!/bin/bash
# TMA tiled-load probe on the GPU box: time + DRAM traffic per L2 promotion mode, against cp.async .L2::64B
o=gpurun_out/${1:-r2o}_tma_probe2.txt
: > $o
B=tools/bin/tma_probe
for rep in 1 2; do
for args in "0 4096 20 1 1 1" "1 4096 20 0 1 1" "1 4096 20 0 0 1" "0 8192 20 1 1 1" "1 8192 20 0 1 1" "1 8192 20 0 0 1"; do
  timeout 120 $B $args >> $o 2>&1
done
done
for args in "0 4096 1 1 1 1" "1 4096 1 0 1 1"; do
  echo "== ncu $args" >> $o
  timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_op_read.sum --clock-control none --launch-skip 2 --launch-count 1 $B $args 2>&1 | grep -E "dram__|gpu__time|lts__|mode" >> $o
done
cat $o
