mkdir -p gpurun_out
set -x
nvidia-smi -L
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_generic.json 2> gpurun_out/bench_generic.err; tail -c 3000 gpurun_out/bench_generic.json; tail -5 gpurun_out/bench_generic.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_generic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_generic.log 2>&1; tail -3 gpurun_out/ncu_generic.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_parity.py -m gpu -q -k "lowlevel or shapes_vs_oracle or t1 or ext_modes" 2>&1 | tail -8
