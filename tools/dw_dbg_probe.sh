python -m pytest tests/test_kernels_gpu.py -x -q -k "dwconv" 2>&1 | tail -3
python tools/kernel_probe.py dwconv --batch 256 --reps 10 2>&1
python tools/kernel_probe.py dwconv --batch 512 --reps 10 2>&1
