# usage: bash tools/gpu/exp23.sh TAG "ENVS run 1" ...  ("-" = no env): GPU parity subset, then one short 2-D and one 3-D bench per env set
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused2d.py tests/test_parity.py tests/test_fullsize.py -m gpu -x -q 2>&1 | tail -4
i=0
for envs in "$@"; do
  i=$((i+1))
  [ "$envs" = "-" ] && envs="DTCWT_B200_NOP=1"
  echo "== run $i: $envs"
  env $envs timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  env $envs timeout 600 python bench.py --workload 3d --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_3d_$i.json 2> gpurun_out/bench_${TAG}_3d_$i.err
  python - <<PY
import json
for f in ("gpurun_out/bench_${TAG}_$i.json", "gpurun_out/bench_${TAG}_3d_$i.json"):
    try:
        d=json.load(open(f))
        print(d["unit"], d["value"], "ms/step", d["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["roofline"]["kernels_ms_per_step"], d["parity"])
    except Exception as e:
        print("bench failed", f, e)
PY
done
