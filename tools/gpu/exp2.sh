# usage: bash tools/gpu/exp2.sh TAG "ENV1=.. ENV2=.." ["ENVS for run 2" ...]
# quick GPU parity subset, then one short 2-D bench per env set ("-" = no env), then one 3-D bench with the first env set
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused2d.py tests/test_parity.py tests/test_fullsize.py -m gpu -x -q 2>&1 | tail -4
i=0
for envs in "$@"; do
  i=$((i+1))
  [ "$envs" = "-" ] && envs="DTCWT_B200_NOP=1"
  echo "== run $i: $envs"
  env $envs timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$i.json"))
    print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["hbm_frac_of_measured_peak"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    print(d["roofline"]["kernels_ms_per_step"], d["parity"]["ok"], d["parity"]["vs_reference_max_rel_err"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${TAG}_$i.err").read()[-1500:])
PY
done
timeout 600 python bench.py --workload 3d --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_3d.json 2> gpurun_out/bench_${TAG}_3d.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_3d.json"))
    print("3d value", d["value"], "ms/step", d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["kernels_ms_per_step"], d["parity"]["ok"])
except Exception as e:
    print("3d bench failed", e); print(open("gpurun_out/bench_${TAG}_3d.err").read()[-1500:])
PY
