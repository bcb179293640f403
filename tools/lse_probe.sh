#!/usr/bin/env bash
# Quick GPU loop for the LSE kernels: time them (sweeping threads per row and rows per
# CTA), then one ncu --set full capture of each.   Usage: bash tools/lse_probe.sh <tag>
set -u
tag=${1:-probe}
out=gpurun_out
mkdir -p $out
: > $out/${tag}_probe.log
for tpr in 64 128 256; do for rpc in 1 2; do
  echo "threads_per_row=$tpr rows_per_cta=$rpc" >> $out/${tag}_probe.log
  B9GW_LSE_THREADS_PER_ROW=$tpr B9GW_LSE_ROWS_PER_CTA=$rpc timeout 300 python tools/lse_probe.py 10000 1024 20 >> $out/${tag}_probe.log 2>&1
done; done
timeout 300 python tools/lse_probe.py 10000 1024 2 > $out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lse_' -s 3 -c 1 \
    -o $out/${tag}_gen python tools/lse_probe.py 10000 1024 2 > $out/${tag}_ncu_gen.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lse_' -s 8 -c 1 \
    -o $out/${tag}_mat python tools/lse_probe.py 10000 1024 2 > $out/${tag}_ncu_mat.log 2>&1
echo "exit=$?" >> $out/${tag}_probe.log
cat $out/${tag}_probe.log
