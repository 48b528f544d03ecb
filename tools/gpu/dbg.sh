mkdir -p gpurun_out
echo "== NO_TMA"; DTCWT_B200_NO_TMA=1 timeout 300 python -m pytest tests/test_fused2d.py -m gpu -x -q 2>&1 | tail -4
echo "== TMA under sanitizer"; timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_fused2d.py -m gpu -x -q -k "shape0" 2>&1 | grep -v "^$" | head -40
