set -u
mkdir -p gpurun_out
python scripts/trace_small.py > gpurun_out/trace_r02d_on.txt 2>&1
FRIEDA_MERKLE_LATENCY=0 python scripts/trace_small.py > gpurun_out/trace_r02d_off.txt 2>&1
python benches/run.py --seconds 0.5 --no-cpu > gpurun_out/criterion_r02d_on.txt 2>&1
FRIEDA_MERKLE_LATENCY=0 python benches/run.py --seconds 0.5 --no-cpu > gpurun_out/criterion_r02d_off.txt 2>&1
cat gpurun_out/trace_r02d_on.txt; echo ----; cat gpurun_out/criterion_r02d_on.txt; echo; echo ---- off; cat gpurun_out/criterion_r02d_off.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-passes --no-extras 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['int_roofline']['frac'])"
