# usage: bash tools/gpu/exp.sh TAG "ENV1=.. ENV2=.." ["ENVS for run 2" ...]   -- GPU tests once, then one short bench per env set
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
i=0
for envs in "$@"; do
  i=$((i+1))
  echo "== run $i: $envs"
  env $envs timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 8 > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$i.json"))
    print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["hbm_frac_of_measured_peak"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    print(d["roofline"]["kernels_ms_per_step"], d["parity"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${TAG}_$i.err").read()[-1500:])
PY
done
