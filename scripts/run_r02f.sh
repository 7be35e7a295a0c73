set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
tail -n 8 gpurun_out/tests.log
timeout 300 python scripts/latency_probe.py > gpurun_out/latency_r02f.txt 2>&1; head -c 2500 gpurun_out/latency_r02f.txt
python benches/run.py --seconds 0.5 --no-cpu 2>&1 | head -22 > gpurun_out/criterion_r02f.txt; cat gpurun_out/criterion_r02f.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-passes 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['int_roofline']['frac'],d['single_blob_latency_ms'],d['prove_c4'],d['c5'])"
