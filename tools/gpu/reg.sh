# usage: bash tools/gpu/reg.sh TAG  -- registration GPU tests + bench line + launch list
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_registration.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload reg --no-e2e > gpurun_out/bench_${TAG}_reg.json 2> gpurun_out/bench_${TAG}_reg.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_reg.json")); print("reg", d["value"], d["ms_per_step"], d["parity"], d["cpu_baseline"]["value"])
PY
bash tools/gpu/listreg.sh $TAG | tail -14
