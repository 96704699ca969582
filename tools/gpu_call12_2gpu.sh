#!/bin/bash
# round 2: the N>1 bench process must leave with rc 0 (teardown order), checked at N=2
set -x
mkdir -p gpurun_out/r02k
O=gpurun_out/r02k
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"
tail -c 600 $O/bench_n2.err; cut -c1-300 $O/bench_n2.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 5 --warmup 3 --workload fifo-da-50k > $O/bench_n2_fifo.json 2> $O/bench_n2_fifo.err; echo "n2 fifo rc=$?"
tail -c 300 $O/bench_n2_fifo.err
