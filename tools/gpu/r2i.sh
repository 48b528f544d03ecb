# round 2, call I: packed row pass of the q-shift inverse; 3-D with two-path loops
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
for cfg in "DTCWT_B200_INVQ_PACKED=1" "DTCWT_B200_INVQ_PACKED=0"; do
  env $cfg timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2i.json"))
    print("$cfg value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], d["parity"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_r2i.err").read()[-2000:])
PY
done
for cfg in "DTCWT_B200_INVQ_PACKED=1" "DTCWT_B200_INVQ_PACKED=0"; do
env $cfg timeout 600 python bench.py --workload 3d --no-cpu-baseline --no-e2e > gpurun_out/bench_r2i_3d.json 2> gpurun_out/bench_r2i_3d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2i_3d.json')); print('3d $cfg', d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'])"
done
