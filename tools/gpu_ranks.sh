#!/usr/bin/env bash
# One gpurun --gpus N call: the N-rank parity script, then the bench at N ranks.  No ncu here
# (never profile a multi-rank command).  Usage: bash tools/gpu_ranks.sh r02 2
set -u
tag=${1:-r02}
n=${2:-2}
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo_n${n}.txt 2>&1
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" \
        --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
run 29511 tests/multigpu/peer_comm_ranks.py > $out/${tag}_ranks_n${n}.log 2>&1
echo "ranks_exit=$?" >> $out/${tag}_ranks_n${n}.log
run 29512 bench.py --gpus "$n" > $out/${tag}_bench_n${n}.json 2> $out/${tag}_bench_n${n}.err
echo "bench_exit=$?" >> $out/${tag}_bench_n${n}.err
# the same collective driven from plain C++ (no Python, no NCCL), at world 1 and n
g++ -std=c++17 -O2 -Iinclude tests/multigpu/peer_comm_c.cpp -o /tmp/peer_comm_c -Lbase_b200 -lb9_groundwork \
    -Loracle -lb9_groundwork_ref -Wl,-rpath,$PWD/base_b200:$PWD/oracle &&
for w in 1 $n; do timeout 120 /tmp/peer_comm_c $w; done > $out/${tag}_cpp_driver_n${n}.log 2>&1
tail -4 $out/${tag}_ranks_n${n}.log; cat $out/${tag}_cpp_driver_n${n}.log; cat $out/${tag}_bench_n${n}.json; tail -3 $out/${tag}_bench_n${n}.err
