# usage: bash tools/gpu/quick.sh TAG [launch-index-to-skip-to ncu-count]
# GPU tests, a short bench (no CPU baseline, no e2e), and optionally a full ncu capture of our kernels.
TAG=${1:-quick}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["hbm_frac_of_measured_peak"], "clocks", d["clocks"])
print(d["roofline"]["kernels_ms_per_step"])
print(d["parity"])
PY
tail -3 gpurun_out/bench_$TAG.err
if [ -n "$2" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd|inv' -s 8 -c 8 -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log
fi
