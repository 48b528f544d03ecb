# usage: bash tools/gpu/exp3d.sh TAG "ENVS run 1" ["ENVS run 2" ...]  ("-" = no env) -- GPU parity subset, then one 3-D bench per env set
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused2d.py tests/test_parity.py tests/test_fullsize.py -m gpu -x -q 2>&1 | tail -4
i=0
for envs in "$@"; do
  i=$((i+1))
  [ "$envs" = "-" ] && envs="DTCWT_B200_NOP=1"
  echo "== 3-D run $i: $envs"
  env $envs timeout 600 python bench.py --workload 3d --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_3d_$i.json 2> gpurun_out/bench_${TAG}_3d_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_3d_$i.json"))
    print("3d value", d["value"], "ms/step", d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["kernels_ms_per_step"], d["parity"])
except Exception as e:
    print("3d bench failed", e); print(open("gpurun_out/bench_${TAG}_3d_$i.err").read()[-1500:])
PY
done
