# round 2 (second session), final collection: smoke, GPU tests, sanitizer gate, the three bench lines in full, reference arms, launch lists,
# full ncu captures (2-D, 3-D)
mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_fused2d.py tests/test_parity.py -m gpu -q -k "roundtrip or fused3d_levels or bp_families or staged or symmetric or transform3d_vs_reference or chained or row_pair" > gpurun_out/sanitizer_memcheck_r3z.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck_r3z.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_fused2d.py -m gpu -q -k "roundtrip or bp_families or staged or symmetric or chained or row_pair" > gpurun_out/sanitizer_racecheck_r3z.log 2>&1; tail -4 gpurun_out/sanitizer_racecheck_r3z.log
timeout 900 python bench.py > gpurun_out/bench_r3z.json 2> gpurun_out/bench_r3z.err; tail -c 400 gpurun_out/bench_r3z.json; tail -2 gpurun_out/bench_r3z.err
timeout 900 python bench.py --workload 3d > gpurun_out/bench_r3z_3d.json 2> gpurun_out/bench_r3z_3d.err; tail -c 300 gpurun_out/bench_r3z_3d.json
timeout 900 python bench.py --workload reg > gpurun_out/bench_r3z_reg.json 2> gpurun_out/bench_r3z_reg.err; tail -c 300 gpurun_out/bench_r3z_reg.json
timeout 600 python bench.py --biort near_sym_a --qshift qshift_a --no-cpu-baseline --no-e2e > gpurun_out/bench_r3z_defaults.json 2> /dev/null
timeout 600 python bench.py --biort near_sym_b_bp --qshift qshift_b_bp --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_r3z_bp.json 2> /dev/null
timeout 300 python tools/exp/config2_latency.py > gpurun_out/config2_latency_r3z.txt 2>&1; tail -3 gpurun_out/config2_latency_r3z.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r3z.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_list_r3z.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd2d|inv2d|invs1' -s 8 -c 8 -o gpurun_out/prof_r3z -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_full_r3z.log 2>&1; tail -1 gpurun_out/ncu_full_r3z.log
python tools/ncu_summary.py gpurun_out/prof_r3z.ncu-rep > gpurun_out/ncu_summary_r3z.txt 2>&1
python tools/ncu_traffic.py gpurun_out/prof_r3z.ncu-rep 4 4096 gpurun_out/ncu_traffic_r3z.json > gpurun_out/ncu_traffic_r3z.txt 2>&1
for i in 0 1 6 7; do python tools/ncu_sass_mix.py gpurun_out/prof_r3z.ncu-rep $i 14; done > gpurun_out/ncu_sass_mix_r3z.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r3z_3d.csv python bench.py --workload 3d --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_list_r3z_3d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd2d|inv2d|z3|axis' -s 20 -c 10 -o gpurun_out/prof_r3z_3d -f python bench.py --workload 3d --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_full_r3z_3d.log 2>&1; tail -1 gpurun_out/ncu_full_r3z_3d.log
python tools/ncu_summary.py gpurun_out/prof_r3z_3d.ncu-rep > gpurun_out/ncu_summary_r3z_3d.txt 2>&1
rm -f gpurun_out/prof_r3z_3d.ncu-rep      # gpurun brings back at most 64 MiB: the 3-D report stays on the box, its summary travels
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_r3z_ref.json 2> /dev/null; cut -c1-300 gpurun_out/bench_r3z_ref.json
du -sh gpurun_out | tail -1
