set -u
mkdir -p gpurun_out
for mode in default chunk10 off; do
  export FRIEDA_MERKLE_LATENCY=1 FRIEDA_MERKLE_LATENCY_CHUNK=8
  [ $mode = chunk10 ] && export FRIEDA_MERKLE_LATENCY_CHUNK=10
  [ $mode = off ] && export FRIEDA_MERKLE_LATENCY=0
  timeout 900 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/tests_$mode.log 2>&1
  echo "== $mode: $(tail -n 1 gpurun_out/tests_$mode.log)"; grep -c FAILED gpurun_out/tests_$mode.log; grep FAILED gpurun_out/tests_$mode.log | head -12 | cut -c1-150
done
