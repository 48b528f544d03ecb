# Round-end style validation: smoke, GPU tests, full bench line, ncu launch list, full ncu capture of our kernels
TAG=${1:-full}
mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 4000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_list_$TAG.log 2>&1; tail -2 gpurun_out/ncu_list_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd|inv' -s 8 -c 8 -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_full_$TAG.log 2>&1; tail -2 gpurun_out/ncu_full_$TAG.log
