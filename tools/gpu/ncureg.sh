# usage: bash tools/gpu/ncureg.sh TAG  -- reg bench line + full ncu capture of the registration kernels (summary only travels back)
TAG=$1
mkdir -p gpurun_out
timeout 600 python bench.py --workload reg --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_reg.json 2> gpurun_out/bench_${TAG}_reg.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_reg.json")); print("reg", d["value"], d["ms_per_step"], d["parity"])
PY
timeout 900 ncu --set full --clock-control none -k regex:generic_1d_kernel -s 40 -c 24 -o gpurun_out/prof_${TAG}_reg -f python bench.py --workload reg --steps 1 --warmup 2 --no-cpu-baseline --no-e2e --images 8 > gpurun_out/ncu_full_${TAG}_reg.log 2>&1; tail -1 gpurun_out/ncu_full_${TAG}_reg.log
python tools/ncu_summary.py gpurun_out/prof_${TAG}_reg.ncu-rep > gpurun_out/ncu_summary_${TAG}_reg.txt 2>&1
rm -f gpurun_out/prof_${TAG}_reg.ncu-rep
cut -c1-900 gpurun_out/ncu_summary_${TAG}_reg.txt
