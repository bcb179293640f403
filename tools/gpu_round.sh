#!/usr/bin/env bash
# One gpurun call on ONE GPU: GPU tests, bench line, ncu launch list, ncu --set full of the
# LSE kernels (memory-fed and register-fed) and the world-1 step kernel.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh r02b
set -u
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/${tag}_gpu.csv
timeout 900 python -m pytest tests -m gpu -x -q -s > $out/${tag}_gpu_tests.log 2>&1
echo "pytest_exit=$?" >> $out/${tag}_gpu_tests.log
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench_exit=$?" >> $out/${tag}_bench.err
timeout 300 python bench.py --impl reference > $out/${tag}_bench_reference.json 2>&1
# design inputs quoted in profiles/: FP64 latency vs chains in flight, LSE time vs problem size
timeout 120 python tools/fp64_latency_probe.py > $out/${tag}_fp64_latency.log 2>&1
for rc in "64 32" "1250 1024" "2500 1024" "5000 1024" "10000 1024" "20000 1024" "40000 1024" "160000 1024" "10000 256" "10000 512"; do
  timeout 120 python tools/lse_probe.py $rc 20
done > $out/${tag}_lse_sizes.log 2>&1
# ncu only after the identical command exited 0 without it
timeout 300 python bench.py --steps 3 --warmup 3 > $out/${tag}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'dfma|rate_kernel|lse_|shard|generate' -c 400 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 > $out/${tag}_ncu_launches.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lse_staged|lse_stream|vshard_step|shard_partials' -c 28 \
    -o $out/${tag}_lse python bench.py --steps 3 --warmup 3 > $out/${tag}_ncu_lse.log 2>&1
echo "ncu_chain_exit=$?" > $out/${tag}_ncu_exit.txt
# the fused step kernel (lse_staged_kernel<1, true>): launches 37-42 of the bench's lse_staged
# launches at --steps 3 --warmup 3 are the 16-chain fused steps (24 plain + 6 share-alone + 6 two-launch first)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lse_staged' -s 38 -c 2 \
    -o $out/${tag}_fused python bench.py --steps 3 --warmup 3 > $out/${tag}_ncu_fused.log 2>&1
echo "ncu_fused_exit=$?" >> $out/${tag}_ncu_exit.txt
tail -5 $out/${tag}_gpu_tests.log; cat $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err; cat $out/${tag}_ncu_exit.txt
