#!/bin/bash
# round 2, GPU call 3 (2 GPUs): the multi-GPU paths -- bench.py under torchrun, gp_multi tests, tools/multi_bench.py
set -x
mkdir -p gpurun_out/r02c
O=gpurun_out/r02c
nvidia-smi -L > $O/smi.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/pytest_multi.txt 2>&1; echo "rc=$?" >> $O/pytest_multi.txt; tail -8 $O/pytest_multi.txt
timeout 300 python bench.py --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "n1 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"; tail -c 1500 $O/bench_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --workload fifo-da-50k > $O/bench_n2_fifo-da-50k.json 2> $O/bench_n2_fifo.err; echo "n2 fifo rc=$?"
timeout 300 python tools/multi_bench.py --config 3 --devices 1 > $O/multi_c3_1.json 2> $O/multi_c3_1.err
timeout 300 python tools/multi_bench.py --config 3 --devices 2 > $O/multi_c3_2.json 2> $O/multi_c3_2.err; tail -c 500 $O/multi_c3_2.err
timeout 300 python tools/multi_bench.py --config 4 --devices 1 > $O/multi_c4_1.json 2> $O/multi_c4_1.err
timeout 300 python tools/multi_bench.py --config 4 --devices 2 > $O/multi_c4_2.json 2> $O/multi_c4_2.err; tail -c 500 $O/multi_c4_2.err
cat $O/*.json | cut -c1-400
