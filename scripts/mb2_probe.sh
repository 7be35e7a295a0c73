set -u
TAG="v1 base" FRIEDA_MERKLE_VARIANT=1 python scripts/merkle_probe.py 1024 2>&1 | tail -1
TAG="v2 mb5" FRIEDA_MERKLE_VARIANT=2 python scripts/merkle_probe.py 1024 2>&1 | tail -1
FRIEDA_MERKLE_VARIANT=2 timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "not multi_gpu" 2>&1 | tail -3
for mb in 4 6; do
  FRIEDA_NVCC_FLAGS="-DFRIEDA_MB2_MIN_BLOCKS=$mb" python -c "from frieda_b200 import build as fb; fb.build(force=True)" > /dev/null 2>&1
  TAG="v2 mb$mb" FRIEDA_MERKLE_VARIANT=2 python scripts/merkle_probe.py 1024 2>&1 | tail -1
done
