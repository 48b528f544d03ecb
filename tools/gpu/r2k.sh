# round 2, call K: GPU tests (bp, keypoints, ...), 2-D and 3-D bench after the latest changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2k.json')); print('2d', d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
timeout 600 python bench.py --workload 3d --no-cpu-baseline --no-e2e > gpurun_out/bench_r2k_3d.json 2> gpurun_out/bench_r2k_3d.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2k_3d.json')); print('3d', d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'])"
timeout 600 python bench.py --biort near_sym_b_bp --qshift qshift_b_bp --no-cpu-baseline --no-e2e --steps 10 > gpurun_out/bench_r2k_bp.json 2> gpurun_out/bench_r2k_bp.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2k_bp.json')); print('bp', d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'])"
