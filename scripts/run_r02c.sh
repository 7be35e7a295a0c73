set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
tail -n 8 gpurun_out/tests.log
timeout 300 python scripts/latency_probe.py > gpurun_out/latency_r02c.txt 2>&1; head -c 2500 gpurun_out/latency_r02c.txt
FRIEDA_MERKLE_LATENCY=0 timeout 300 python scripts/latency_probe.py > gpurun_out/latency_r02c_off.txt 2>&1; head -n 4 gpurun_out/latency_r02c_off.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo "bench exit $?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_r02c.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['int_roofline']['frac'], d.get('single_blob_latency_ms'), d['prove_c4'], d['c5'])
print(d.get('kernels'))
P
