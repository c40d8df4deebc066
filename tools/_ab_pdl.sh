set -x
mkdir -p gpurun_out
for p in 0 1; do
  echo "=== HPVG_PDL=$p" 
  HPVG_PDL=$p timeout 300 python tools/train_vae_only.py 200 20 graph
  HPVG_PDL=$p timeout 300 python tools/train_vae_only.py 200 20 graph 2d
  HPVG_PDL=$p timeout 300 python tools/train_only.py 20 5 graph
  HPVG_PDL=$p timeout 300 python bench.py --steps 10 --warmup 3 --no-extras
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
