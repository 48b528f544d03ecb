# round 2, call F: round-end style collection -- smoke, GPU tests, the three bench lines in full (CPU baseline, e2e), the
# reference arm, launch lists and full ncu captures for the 2-D and 3-D workloads
mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -c 600 gpurun_out/bench_r2f.json; tail -3 gpurun_out/bench_r2f.err
timeout 900 python bench.py --workload 3d > gpurun_out/bench_r2f_3d.json 2> gpurun_out/bench_r2f_3d.err; tail -c 300 gpurun_out/bench_r2f_3d.json; tail -3 gpurun_out/bench_r2f_3d.err
timeout 900 python bench.py --workload reg > gpurun_out/bench_r2f_reg.json 2> gpurun_out/bench_r2f_reg.err; tail -c 1500 gpurun_out/bench_r2f_reg.json; tail -3 gpurun_out/bench_r2f_reg.err
timeout 600 python bench.py --biort near_sym_a --qshift qshift_a --no-cpu-baseline --no-e2e > gpurun_out/bench_r2f_defaults.json 2> /dev/null; tail -c 300 gpurun_out/bench_r2f_defaults.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r2f.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_list_r2f.log 2>&1; tail -1 gpurun_out/ncu_list_r2f.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd2d|inv2d|invs1' -s 8 -c 8 -o gpurun_out/prof_r2f -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_full_r2f.log 2>&1; tail -1 gpurun_out/ncu_full_r2f.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2f_3d.csv python bench.py --workload 3d --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_list_r2f_3d.log 2>&1; tail -1 gpurun_out/ncu_list_r2f_3d.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd2d|inv2d|z3|axis' -s 20 -c 10 -o gpurun_out/prof_r2f_3d -f python bench.py --workload 3d --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_full_r2f_3d.log 2>&1; tail -1 gpurun_out/ncu_full_r2f_3d.log
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 --workload 3d > gpurun_out/bench_r2f_ref3d.json 2> /dev/null; cat gpurun_out/bench_r2f_ref3d.json | cut -c1-400
