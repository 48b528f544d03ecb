# round 2, call J (8 GPUs): bare pinned-copy bandwidth with all GPUs copying at once, the bench line at N=8 (weak), and the
# 1024-image job of SURVEY 8(d)/(e) (strong)
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
for n in 8 4 2; do timeout 300 python tools/exp/copy_bw.py --gpus $n > gpurun_out/copy_bw_$n.json 2> gpurun_out/copy_bw_$n.err; cat gpurun_out/copy_bw_$n.json; done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 > gpurun_out/bench_r2j_8gpu.json 2> gpurun_out/bench_r2j_8gpu.err; tail -c 2500 gpurun_out/bench_r2j_8gpu.json; tail -3 gpurun_out/bench_r2j_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --total-images 1024 --no-e2e > gpurun_out/bench_r2j_8gpu_strong.json 2> gpurun_out/bench_r2j_8gpu_strong.err; tail -c 1200 gpurun_out/bench_r2j_8gpu_strong.json; tail -3 gpurun_out/bench_r2j_8gpu_strong.err
