#!/bin/bash
# Run on the GPU box (gpurun): the round's evidence in one call.  Outputs under gpurun_out/evidence/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/collect_evidence.sh r02'
set -u
exec < /dev/null
TAG=${1:-r02}
OUT=gpurun_out/evidence
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/${TAG}_gpu.csv
# 1. parity: the whole GPU suite
(timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_pytest_gpu.log)
tail -3 $OUT/${TAG}_pytest_gpu.log
# 2. smoke
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
# 3. the bench line (N = 1) and the reference arm
timeout -k 10 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; tail -2 $OUT/${TAG}_bench_n1.err
timeout -k 10 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench_n1.err
# 4. launch list of the same command (ncu per-launch times are cold-cache and serialised: compare SHARES)
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"osc_step|pack_tiles|calc_error" -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --no-graph --no-cpu-baseline --no-extras --steps 4 --warmup 3 > $OUT/${TAG}_launches_run.log 2>&1
# 5. one full capture of the default kernel per workload
for wl in gain_test admit_test worst_case; do
    timeout -k 10 300 ncu --set full --clock-control none --import-source on -k "regex:osc_step_(lane|pair)" -s 8 -c 1 -f -o $OUT/${TAG}_tiles_$wl \
        python bench.py --workload $wl --no-graph --no-extras --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > $OUT/${TAG}_ncu_$wl.log 2>&1
done
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:osc_step_tree -s 3 -c 1 -f -o $OUT/${TAG}_tree_gain_test \
    python tools/lane_bench.py --scenario gain_test --threads 224 --staged 1 --stages 3 --others tree_qm --iters 5 > $OUT/${TAG}_ncu_tree.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:osc_step_fused -s 3 -c 1 -f -o $OUT/${TAG}_fused_gain_test \
    python tools/lane_bench.py --scenario gain_test --threads 224 --staged 1 --stages 3 --others fused --iters 5 > $OUT/${TAG}_ncu_fused.log 2>&1
# summaries + traffic.json from the reports, then drop all but one report (gpurun brings back at most 64 MiB)
timeout -k 10 300 python tools/evidence_to_profiles.py $TAG --on-box > $OUT/${TAG}_summarise.log 2>&1; tail -3 $OUT/${TAG}_summarise.log
ls -la $OUT; du -sh $OUT
