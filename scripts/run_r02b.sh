set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/tests.log
tail -n 15 gpurun_out/tests.log
for n in 512 2048; do timeout 300 python scripts/prove_probe.py $n > gpurun_out/p${n}_r02b.log 2>&1; tail -n 4 gpurun_out/p${n}_r02b.log; done
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_r02b.json
for L in 2 4; do FRIEDA_MERKLE_LEVELS=$L timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-passes --no-extras > gpurun_out/bench_levels${L}.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/bench_levels${L}.json').read().strip().splitlines()[-1]);print('levels',${L},d['value'],d['e2e']['value'])"; done
