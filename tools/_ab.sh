python tools/perf_conv.py 2>&1 | tail -14
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_tf32.py tests/test_gpu_backward.py -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-extras | cut -c1-200
timeout 300 python tools/train_only.py 20 5 graph | cut -c1-200
timeout 300 python tools/train_vae_only.py 200 20 graph | cut -c1-200
