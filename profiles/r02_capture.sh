#!/bin/bash
# How the round-2 evidence under profiles/ was captured (one gpurun call on the final build; outputs land in gpurun_out/ and are
# summarised with summarize.sh, ncu_source.py, sass_excerpt.py): full GPU tests, the default bench line, launch lists,
# ncu --set full captures of one cfg2 / cfg3 / cfg4 frame, the timeline trace.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r5_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r5_tests.log
timeout 600 python bench.py > gpurun_out/r5_bench_default.json 2> gpurun_out/r5_bench_default.err
B="--steps 4 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-other-configs"
for c in cfg2 cfg3 cfg4; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r5_${c}_launches.csv python bench.py --config $c $B > gpurun_out/r5_${c}_launches.out 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gatherFrameKernel -s 5 -c 1 -f -o gpurun_out/r5_cfg2_frame python bench.py --config cfg2 $B > gpurun_out/r5_ncu_cfg2.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:blurFrameStripKernel|gatherFrameKernel' -s 6 -c 2 -f -o gpurun_out/r5_cfg3_frame python bench.py --config cfg3 $B > gpurun_out/r5_ncu_cfg3.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:blurFrameStripKernel|gatherFrameKernel' -s 6 -c 2 -f -o gpurun_out/r5_cfg4_frame python bench.py --config cfg4 $B > gpurun_out/r5_ncu_cfg4.out 2>&1
timeout 200 python profiles/trace_frame.py cfg2 > gpurun_out/r5_trace_cfg2.txt 2>&1
tail -3 gpurun_out/r5_tests.log
tail -c 3000 gpurun_out/r5_bench_default.json
ls -la gpurun_out/r5_*
cat gpurun_out/r5_trace_cfg2.txt
