#!/bin/bash
set -x
mkdir -p gpurun_out/r02g
O=gpurun_out/r02g
python -c "import __graft_entry__ as g; g.build()" > $O/build.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "rc=$?" >> $O/pytest_gpu.txt; tail -12 $O/pytest_gpu.txt | cut -c1-250
for w in fifo-10k fifo-da-50k; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"
done
GANGPACK_TRACE=2 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_trace2.json 2> $O/bench_trace2.err
GANGPACK_TRACE=2 GANGPACK_CHUNK_APPS=34000 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_trace2_34k.json 2> $O/bench_trace2_34k.err
python tools/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1
# scan path (node-order scan kernel) on the headline workload: ncu capture for the roofline object
GANGPACK_TABLES=0 BENCH_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_pack_listed -s 3 -c 1 -o $O/r02_listed_scan python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_listed.log 2>&1
BENCH_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_build_shape_tables -s 3 -c 1 -o $O/r02_tables python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_tables.log 2>&1
BENCH_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_pack_fifo_cta -s 2 -c 1 -o $O/r02_fifo_v4 python bench.py --workload fifo-10k --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_fifo.log 2>&1
# sanitizer (time-boxed): memcheck over the small-input parity tests of every kernel family, racecheck over the shared-memory kernels
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zones.py tests/test_gpu_wire_and_tables.py tests/test_gpu_reference_scenarios.py -m gpu -q -x \
   -k "golden or random_independent or random_fifo or multi_group or awkward or degenerate or error_paths or reschedule or potential_nodes_random or single_az or reservation_table or rejections or tables_multi_group or undefined or scenarios or minimal_fragmentation" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.log; tail -5 $O/sanitizer_memcheck.log
timeout 480 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wire_and_tables.py -m gpu -q -x \
   -k "golden_fifo or (random_fifo and 0-) or multi_group or tables_multi_group or potential_nodes_goldens" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.log; tail -5 $O/sanitizer_racecheck.log
ls $O
