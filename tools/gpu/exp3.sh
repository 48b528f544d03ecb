# usage: bash tools/gpu/exp3.sh TAG   -- chained levels 1+2 through L2: parity test, bench per mode, DRAM bytes per launch
TAG=$1
mkdir -p gpurun_out
python - <<'PY'
import ctypes
from dtcwt_b200 import _lib
import torch
torch.zeros(1, device="cuda")
out = (ctypes.c_int64 * 3)()
_lib.lib().dtcwt_b200_l2_info(out)
print("l2_info: max persisting %.1f MiB, max window %.1f MiB, L2 %.1f MiB" % tuple(v / 2**20 for v in out))
PY
timeout 900 python -m pytest tests/test_fused2d.py -m gpu -x -q 2>&1 | tail -3
i=0
for envs in "DTCWT_B200_CHAIN=0" "DTCWT_B200_CHAIN=-1" "DTCWT_B200_CHAIN=100" "DTCWT_B200_CHAIN=60" "DTCWT_B200_CHAIN=100 DTCWT_B200_CHAIN_MB=128"; do
  i=$((i+1))
  echo "== run $i: $envs"
  env $envs timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$i.json"))
    print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["hbm_frac_of_measured_peak"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "launches", d["gpu_launches"])
    print(d["roofline"]["kernels_ms_per_step"], d["parity"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_${TAG}_$i.err").read()[-1500:])
PY
done
for mode in 0 100; do
  DTCWT_B200_CHAIN=$mode timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --cache-control none -k regex:'fwd2d|inv2d|invs1' -c 80 --csv --log-file gpurun_out/dram_${TAG}_chain$mode.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --images 4 > gpurun_out/ncu_dram_${TAG}_$mode.log 2>&1
  tail -2 gpurun_out/ncu_dram_${TAG}_$mode.log
done
