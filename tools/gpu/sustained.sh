# usage: bash tools/gpu/sustained.sh TAG "ENVS 1" "ENVS 2" ...  -- long (sustained-power) 2-D bench per env set, alternating twice
TAG=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
  i=0
  for envs in "$@"; do
    i=$((i+1))
    [ "$envs" = "-" ] && envs="DTCWT_B200_NOP=1"
    env $envs timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 80 --warmup 20 > gpurun_out/bench_${TAG}_${i}_$rep.json 2> /dev/null
    python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_${i}_$rep.json"))
c=d["clocks"]
print("rep $rep [$envs]", d["value"], "ms/step", d["ms_per_step"], "clk", c["sm_mhz"], c.get("sm_min_mhz"), c["reasons"], "W", c.get("power_w_max"), "dominant avg", d["roofline"]["avg_launch_ms"])
PY
  done
done
nvidia-smi --query-gpu=power.limit,power.max_limit,power.default_limit,clocks.max.sm --format=csv
