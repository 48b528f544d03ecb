# usage: bash tools/gpu/ncu1.sh TAG "ENVS"   -- one full ncu capture of our 8 kernels (4 images), no tests/bench
TAG=$1
mkdir -p gpurun_out
env $2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd|inv' -s 8 -c 8 -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log
