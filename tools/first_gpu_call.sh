#!/usr/bin/env bash
# First gpurun call of a round: everything that was written without a GPU gets its first run, and the numbers
# needed to decide the next kernel change come back in gpurun_out/.
#
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
#
# 1. GPU test suite (the test_gpu_zz_* files are the ones that have never run on a GPU)
# 2. bench line (its "side_legs" key carries qM vs packed, iros2022, the 12-action sequence, cooperative fix-up)
# 3. qM vs packed for the streaming kernel at two batch sizes, then the launch list of a short bench run
# 4. one ncu --set full capture of the streaming kernel reading qM
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/t_first.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_first.log
tail -3 gpurun_out/t_first.log
python bench.py > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench rc=$?"
for B in 65536 262144; do
  for L in packed qM; do
    python bench.py --kernel 9 --m-layout $L --batch $B --no-side-legs --no-fused --no-cpu-baseline --no-e2e \
      > gpurun_out/stream_${L}_B${B}.json 2>> gpurun_out/bench_first.err
  done
done
for K in 2 5 6; do      # tree-sparse kernel on qM: table variants 0 (w8 s4), 3 (w7 s4), 4 (w8 s3); packed default = --kernel 0
  python bench.py --kernel $K --m-layout qM --no-side-legs --no-fused --no-cpu-baseline --no-e2e \
    > gpurun_out/tree_qM_k${K}.json 2>> gpurun_out/bench_first.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_first.csv \
  python bench.py --steps 4 --warmup 3 --no-side-legs --no-cpu-baseline > gpurun_out/ncu_bench_first.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:osc_step_stream -s 3 -c 1 \
  -o gpurun_out/stream_qM python bench.py --kernel 9 --m-layout qM --steps 2 --warmup 3 --no-side-legs --no-fused \
  --no-cpu-baseline --no-e2e > gpurun_out/ncu_stream_qM.log 2>&1
ls -la gpurun_out | tail -12
