#!/bin/bash
# Round-end GPU record (one B200): bench line, ncu launch list of the same command, `ncu --set full` captures of the
# dominant kernels, criterion table, single-blob latency, compute-sanitizer passes.  Run under gpurun from the repo root:
#   gpurun --timeout 2400 -- 'bash scripts/final_profile.sh r02'      (SKIP_NCU=1: everything but the ncu passes;
#   the .ncu-rep files of one full run come to ~70 MB, above what gpurun merges back: fetch them in two calls)
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
# 1. the bench line (never under a profiler)
timeout 900 python bench.py --steps 3 --warmup 3 > $O/${R}_bench_1gpu.json 2> $O/${R}_bench_1gpu.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference.json 2>/dev/null; echo "reference arm exit $?"
if [ -z "${SKIP_NCU:-}" ]; then
# 2. launch list of the bench command (per-launch times are cold-cache and serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv \
  python bench.py --steps 1 --warmup 3 --blobs 512 --no-cpu-baseline --no-passes > $O/ncu_bench.log 2>&1; echo "ncu list exit $?"
# 3. full captures: second FRI-commit wave of 296 blobs (the first is warm-up; 16 merkle_bottom launches per wave:
#    columns, circle fold, 7 line folds, 7 middle passes), the LDE pass, one grind launch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:merkle_bottom_kernel -s 16 -c 16 -f \
  -o $O/${R}_merkle_bottom python scripts/prof_target.py 296 > $O/ncu_m.log 2>&1; echo "ncu merkle exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lde_warp_kernel -s 1 -c 1 -f \
  -o $O/${R}_lde_c2 python scripts/prof_target.py 256 > $O/ncu_l.log 2>&1; echo "ncu lde exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grind_kernel -s 4 -c 1 -f \
  -o $O/${R}_grind python scripts/prove_probe.py 512 > $O/ncu_g.log 2>&1; echo "ncu grind exit $?"
python scripts/ncu_summary.py $O/${R}_merkle_bottom.ncu-rep $O/${R}_lde_c2.ncu-rep $O/${R}_grind.ncu-rep > $O/${R}_ncu_metrics.txt 2>&1
rm -f $O/${R}_ncu_metrics.json
python scripts/ncu_summary.py --json $O/${R}_ncu_metrics.json --blobs 296 --git "$(cat $O/../.git_head 2>/dev/null || echo unknown)" \
  $O/${R}_merkle_bottom.ncu-rep > $O/ncu_json.log 2>&1
fi
# 4. the reference's criterion groups and single-call latencies
timeout 600 python benches/run.py --seconds 0.5 > $O/${R}_criterion.txt 2>&1; echo "criterion exit $?"
timeout 300 python scripts/latency_probe.py > $O/${R}_latency.txt 2>&1
for n in 512 2048; do timeout 300 python scripts/prove_probe.py $n; done > $O/${R}_prove_probe.txt 2>&1
# 5. compute-sanitizer on the kernels that changed this session (latency-form Merkle passes, grind grid, pipelined proofs)
K='(merkle_pass_vs_oracle and not 21) or prove_batch_in_waves or proof_edge or test_verify_proof or (fri_commit_every_intermediate and pattern-3000)'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$K" > $O/${R}_san_mem.log 2>&1; echo "memcheck exit $?"
K2='(merkle_pass_vs_oracle and (10 or 11 or 13)) or (fri_commit_every_intermediate and pattern-3000) or (proof_edge and 6000)'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$K2" > $O/${R}_san_race.log 2>&1; echo "racecheck exit $?"
tail -n 3 $O/${R}_san_mem.log $O/${R}_san_race.log
ls -la $O/${R}_*
tail -c 600 $O/${R}_bench_1gpu.json
