# round 2, call D: all GPU tests; 2-D bench with the staged (bulk-copy) level-1 inverse vs the per-thread-load one; ncu of the staged kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
for st in 1 0; do
  DTCWT_B200_INV_STAGED=$st timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_r2d_st$st.json 2> gpurun_out/bench_r2d_st$st.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2d_st$st.json"))
    print("STAGED=$st value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["kernels_ms_per_step"], d["parity"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_r2d_st$st.err").read()[-2000:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'invs1t' -s 2 -c 2 -o gpurun_out/prof_r2d_invs1t -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_r2d.log 2>&1; tail -2 gpurun_out/ncu_r2d.log
timeout 600 python bench.py --workload 3d --no-cpu-baseline --no-e2e > gpurun_out/bench_r2d_3d.json 2> gpurun_out/bench_r2d_3d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2d_3d.json')); print('3d', d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'])"
