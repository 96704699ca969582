#!/bin/bash
# round 2, GPU call 1: state after the evidence fixes (bench teardown, parity word, full-size parity tests)
set -x
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
timeout 300 python bench.py > $O/bench_tightly-100k.json 2> $O/bench_default.err; echo "bench rc=$?"
tail -c 600 $O/bench_default.err
for w in tightly-100k-deep evenly-100k-deep evenly-100k; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>&1
# current minimal-fragmentation kernel (never re-captured after the round-1 rewrite)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_pack_independent -s 3 -c 1 -o $O/r02_minfrag \
   python bench.py --workload minfrag-100k --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_minfrag.log 2>&1
BENCH_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_pack_independent -s 3 -c 1 -o $O/r02_deep_tightly \
   python bench.py --workload tightly-100k-deep --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_deep.log 2>&1
python tools/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1
ls -la $O
