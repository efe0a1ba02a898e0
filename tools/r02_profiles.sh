#!/bin/bash
# Round-2 measurement + profiling artefacts (run on a B200 box through gpurun; outputs under gpurun_out/, the summaries
# are then written to profiles/ on the build container by tools/r02_collect.py).
set -x
O=gpurun_out
# 0. the driver's own checks: GPU test suite, smoke, default bench (both arms)
python -m pytest tests -m gpu -x -q > $O/r02_final_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_final_smoke.log 2>&1
python bench.py > $O/r02_final_bench.json 2> $O/r02_final_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_final_ref.json 2> $O/r02_final_ref.err
# 1. launch list of the bench command (graph replays included: ncu profiles the kernel nodes one by one)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $O/r02_launches_bench.log 2>&1
# 2. one step, every kernel, durations + instruction counts (AG_NO_GRAPH: one launch per kernel, in order)
AG_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --profile-from-start off --csv --log-file $O/r02_step_kernels.csv \
    python tools/profile_step.py 2 3 0 > $O/r02_step_kernels.log 2>&1
# 3. full capture of the kernels that own the step
AG_NO_GRAPH=1 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:"k_hand_sweep|k_hog_svm|k_taubin_solve|k_rank_picks|k_ball_moments|k_emit_bitmap|k_axes_finish|k_export|k_classify|k_mark" \
    -o $O/r02_full -f python tools/profile_step.py 2 3 0 > $O/r02_full.log 2>&1
ncu -i $O/r02_full.ncu-rep --page raw --csv > $O/r02_full_raw.csv 2>/dev/null
for k in k_hand_sweep k_hog_svm k_taubin_solve k_rank_picks k_ball_moments k_emit_bitmap; do
  ncu -i $O/r02_full.ncu-rep --page source --csv --kernel-name regex:$k > $O/r02_full_src_$k.csv 2>/dev/null
done
rm -f $O/r02_full.ncu-rep
# 4. parity statistics at the full BASELINE sizes
python tools/parity_report.py > $O/r02_parity.log 2>&1
