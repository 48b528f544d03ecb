# round 2, call B: fused 3-D levels on hardware: GPU tests, config-4 bench line, launch list + one full ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --workload 3d > gpurun_out/bench_r2b_3d.json 2> gpurun_out/bench_r2b_3d.err; tail -c 4000 gpurun_out/bench_r2b_3d.json; tail -5 gpurun_out/bench_r2b_3d.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2b_3d.csv python bench.py --workload 3d --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 2 > gpurun_out/ncu_list_r2b.log 2>&1; tail -2 gpurun_out/ncu_list_r2b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd2d|inv2d|z3|axis' -s 20 -c 10 -o gpurun_out/prof_r2b_3d -f python bench.py --workload 3d --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 2 > gpurun_out/ncu_full_r2b.log 2>&1; tail -2 gpurun_out/ncu_full_r2b.log
