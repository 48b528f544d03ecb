# Round 1, step 3: fused per-level kernels -- smoke, tests, bench (full), launch list, full ncu capture
mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -5
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; tail -c 3500 gpurun_out/bench_fused.json; tail -5 gpurun_out/bench_fused.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_fused.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_fused.log 2>&1; tail -3 gpurun_out/ncu_fused.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd2d|inv2d' -s 16 -c 8 -o gpurun_out/prof_fused -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
