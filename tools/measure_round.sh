# The single-GPU measurements of a round (tests, bench lines, timelines, ncu launch list and full captures): run on the GPU box,
# results under gpurun_out/; what is kept is copied to profiles/ by hand.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
python bench.py > gpurun_out/r02_bench_c2_n1.json 2> gpurun_out/r02_bench_c2_n1.err; tail -c 300 gpurun_out/r02_bench_c2_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_c2_reference.json 2>/dev/null; tail -c 300 gpurun_out/r02_bench_c2_reference.json
python bench.py --workload c3 --no-configs > gpurun_out/r02_bench_c3_b64.json 2>/dev/null
python bench.py --workload c4 --no-configs > gpurun_out/r02_bench_c4_gray.json 2>/dev/null
python bench.py --family 1 --no-configs > gpurun_out/r02_bench_c2_noise.json 2>/dev/null
python tools/kernel_timeline.py --steps 30 > gpurun_out/r02_timeline_c2.txt 2>&1
python tools/kernel_timeline.py --workload c3 --steps 5 > gpurun_out/r02_timeline_c3.txt 2>&1
python tools/pcie_probe.py > gpurun_out/r02_pcie_probe.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-configs > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fwd_transform|k_inv_transform2|k_sync_decode_hyp|k_write_coefs|k_scatter|k_block_bits" -s 30 -c 6 -o gpurun_out/r02_full -f python bench.py --steps 3 --warmup 3 --no-e2e --no-configs > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | tail -12
