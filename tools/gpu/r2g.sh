# round 2, call G: level-1 forward with shared symmetric sums in the column pass vs the scatter form
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused2d.py -m gpu -q 2>&1 | tail -3
for cfg in "DTCWT_B200_FWD_SYM=1" "DTCWT_B200_FWD_SYM=0"; do
  env $cfg timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2g.json"))
    print("$cfg value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], d["parity"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_r2g.err").read()[-2000:])
PY
done
DTCWT_B200_FWD_SYM=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd2d_kernel.*Li3E' -s 0 -c 1 -o gpurun_out/prof_r2g_sym -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_r2g.log 2>&1; tail -1 gpurun_out/ncu_r2g.log
