# round 2, call C: find the ext_mode-8 3-D failure, run all GPU tests (registration included), 3-D bench with both depth variants
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_parity.py -m gpu -x -q -k "fused3d_levels and shape3" 2>&1 | grep -v "^$" | head -60
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12
for ng in 4 2; do
  DTCWT_B200_Z3_NG=$ng timeout 600 python bench.py --workload 3d --no-cpu-baseline --no-e2e > gpurun_out/bench_r2c_3d_ng$ng.json 2> gpurun_out/bench_r2c_3d_ng$ng.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2c_3d_ng$ng.json"))
print("NG=$ng value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["kernels_ms_per_step"])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2c_3d.csv python bench.py --workload 3d --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_r2c.log 2>&1; tail -2 gpurun_out/ncu_list_r2c.log
