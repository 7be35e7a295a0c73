# usage: bash scripts/run_multi.sh N   (under gpurun --gpus N)
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/tests_multi_${N}.log 2>&1; echo "multi tests exit $?"; tail -n 3 gpurun_out/tests_multi_${N}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench${N}_r02.json 2> gpurun_out/bench${N}_r02.err; echo "bench exit $?"
tail -c 900 gpurun_out/bench${N}_r02.json; tail -n 5 gpurun_out/bench${N}_r02.err
