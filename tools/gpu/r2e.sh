# round 2, call E: staged level-1 inverse after the one-arrival-per-warp fix (depth 4 and 5) vs per-thread loads; ncu of the staged kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
for cfg in "DTCWT_B200_INV_STAGED=1" "DTCWT_B200_INV_STAGED=1 DTCWT_B200_INV_NSTAGE=4" "DTCWT_B200_INV_STAGED=0"; do
  env $cfg timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2e.json"))
    print("$cfg value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_r2e.err").read()[-2000:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'invs1t' -s 2 -c 1 -o gpurun_out/prof_r2e_invs1t -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_r2e.log 2>&1; tail -2 gpurun_out/ncu_r2e.log
