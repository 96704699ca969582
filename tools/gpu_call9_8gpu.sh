#!/bin/bash
# round 2, 8-GPU call: scaling of bench.py (one process per GPU) and of the one-process gp_multi handle
set -x
mkdir -p gpurun_out/r02h
O=gpurun_out/r02h
nvidia-smi -L > $O/smi.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.txt 2>&1
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "n$n rc=$?"; tail -c 400 $O/bench_n$n.err
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 4 --steps 10 --warmup 3 --workload fifo-da-50k > $O/bench_n4_fifo-da-50k.json 2> $O/bench_n4_fifo.err; echo "n4 fifo rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 10 --warmup 3 --workload tightly-50k-1m > $O/bench_n8_tightly-50k-1m.json 2> $O/bench_n8_50k.err; echo "n8 50k rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 8 --steps 2 --warmup 1 --impl reference > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err
for d in 1 2 4 8; do timeout 200 python tools/multi_bench.py --config 4 --devices $d --steps 6 --warmup 2 > $O/multi_c4_$d.json 2> $O/multi_c4_$d.err; done
for d in 1 4; do timeout 200 python tools/multi_bench.py --config 3 --devices $d --steps 6 --warmup 2 > $O/multi_c3_$d.json 2> $O/multi_c3_$d.err; done
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/pytest_multi.txt 2>&1; tail -3 $O/pytest_multi.txt
cat $O/*.json | cut -c1-260
