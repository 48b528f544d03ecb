# round 2, call H: tile-shape variants of the forward q-shift kernel; symmetric-sum forward as default; 3-D with the interior fast path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
for v in 0 1 2 3; do
  DTCWT_B200_FWDQ_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2h.json"))
    print("FWDQ_VARIANT=$v value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_r2h.err").read()[-2000:])
PY
done
timeout 600 python bench.py --workload 3d --no-cpu-baseline --no-e2e > gpurun_out/bench_r2h_3d.json 2> gpurun_out/bench_r2h_3d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2h_3d.json')); print('3d', d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'])"
