mkdir -p gpurun_out
python scripts/gpu_debug.py golden_cfg1_600 cfg1_1500 2>&1 | grep -v "bad bytes 0" | tail -5
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu_r01b.log 2>&1; tail -8 gpurun_out/pytest_gpu_r01b.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
