#!/bin/bash
# round 2, GPU call (1 GPU): full test suite, benches, ncu of the FIFO and decide kernels
set -x
mkdir -p gpurun_out/r02d
O=gpurun_out/r02d
python -c "import __graft_entry__ as g; g.build()" > $O/build.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "rc=$?" >> $O/pytest_gpu.txt; tail -25 $O/pytest_gpu.txt | cut -c1-250
timeout 300 python bench.py > $O/bench_tightly-100k.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 800 $O/bench_default.err
GANGPACK_CHUNK_APPS=25000 timeout 300 python bench.py --no-cpu-baseline > $O/bench_tightly-100k_chunk25k.json 2>/dev/null
GANGPACK_CHUNK_APPS=1000000 timeout 300 python bench.py --no-cpu-baseline > $O/bench_tightly-100k_nochunk.json 2>/dev/null
GANGPACK_TABLES=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_tightly-100k_scan.json 2> $O/bench_scan.err
for w in evenly-100k tightly-100k-deep evenly-100k-deep fifo-10k fifo-da-50k tightly-50k-1m tightly-10k minfrag-100k; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"
done
GANGPACK_TABLES=0 timeout 300 python bench.py --workload evenly-100k-deep --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_evenly-100k-deep_scan.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>&1
GANGPACK_TRACE=1 python tools/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1
BENCH_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches.log 2>&1
BENCH_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_pack_fifo_cta -s 2 -c 1 -o $O/r02_fifo python bench.py --workload fifo-10k --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_fifo.log 2>&1
BENCH_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_decide_tables -s 3 -c 1 -o $O/r02_decide python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_decide.log 2>&1
ls -la $O
