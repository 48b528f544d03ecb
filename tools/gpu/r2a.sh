# round 2, call A: validate phase A on hardware: smoke, GPU tests (incl. push_backend through oracle/_ref), full bench line
# with the real reference as cpu_baseline + full-array parity, the reference arm, bare copy bandwidth
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python tools/exp/copy_bw.py --gpus 1 > gpurun_out/copy_bw_1.json 2> gpurun_out/copy_bw_1.err; cat gpurun_out/copy_bw_1.json
timeout 900 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 5000 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_r2a_ref.json 2> gpurun_out/bench_r2a_ref.err; cat gpurun_out/bench_r2a_ref.json; tail -3 gpurun_out/bench_r2a_ref.err
