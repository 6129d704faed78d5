#!/bin/bash
# round-end measurement set on one B200: tests, smoke, benches (ncu passes: see profiles/r01_summary.md for the commands)
set -u
mkdir -p gpurun_out/final
O=gpurun_out/final
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_c2_default.json 2> $O/bench_c2_default.err
python bench.py --impl reference > $O/bench_c2_reference.json 2> $O/bench_c2_reference.err
python bench.py --workload c3 --steps 20 --warmup 3 > $O/bench_c3_b64.json 2> $O/bench_c3_b64.err
python bench.py --workload c3 --family 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_c3_noise.json 2> $O/bench_c3_noise.err
python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_c4_gray.json 2> $O/bench_c4_gray.err
python bench.py --workload c1 --no-cpu-baseline > $O/bench_c1.json 2> $O/bench_c1.err
python tools/kernel_timeline.py --workload c2 > $O/timeline_c2.txt 2>&1
python tools/kernel_timeline.py --workload c3 --steps 5 > $O/timeline_c3.txt 2>&1
ls -la $O | head -30
