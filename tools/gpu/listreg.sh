# usage: bash tools/gpu/listreg.sh TAG  -- ncu launch list (durations only) of the registration step
TAG=$1
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${TAG}_reg.csv python bench.py --workload reg --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 8 > gpurun_out/ncu_list_${TAG}_reg.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_${TAG}_reg.csv 2>&1 | tail -40
