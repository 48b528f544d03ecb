# usage: bash tools/gpu/list3d.sh TAG  -- ncu launch lists (durations only) of the 3-D and 2-D steps
TAG=$1
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}_3d.csv python bench.py --workload 3d --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_list_${TAG}_3d.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_${TAG}_3d.csv 2>&1 | tail -16
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_list_${TAG}.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_${TAG}.csv 2>&1 | tail -12
