#!/bin/bash
set -x
mkdir -p gpurun_out/r02f
O=gpurun_out/r02f
python -c "import __graft_entry__ as g; g.build()" > $O/build.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "rc=$?" >> $O/pytest_gpu.txt; tail -12 $O/pytest_gpu.txt | cut -c1-250
timeout 300 python bench.py > $O/bench_tightly-100k.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 600 $O/bench_default.err
GANGPACK_GRAPHS=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_tightly-100k_nographs.json 2>/dev/null
GANGPACK_CHUNK_APPS=25000 timeout 300 python bench.py --no-cpu-baseline > $O/bench_tightly-100k_chunk25k.json 2>/dev/null
GANGPACK_CHUNK_APPS=34000 timeout 300 python bench.py --no-cpu-baseline > $O/bench_tightly-100k_chunk34k.json 2>/dev/null
for w in fifo-10k fifo-da-50k evenly-100k; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"
done
GANGPACK_TRACE=1 python tools/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1
BENCH_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_pack_fifo_cta -s 2 -c 1 -o $O/r02_fifo_v3 python bench.py --workload fifo-10k --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_fifo.log 2>&1
ls $O
