#!/bin/bash
# Round evidence on the GPU box: plain bench (the judged numbers), then the ncu launch list of the same command,
# then one `--set full` capture of each kernel of a full-size step.  Usage: tools/profile_round.sh <tag> [steps]
tag=${1:-r1x}; steps=${2:-400}
o=gpurun_out
KF='regex:roi_|ring_push|firls|butter|window_preprocess|spectrum_dense|welch|xcorr|ls_coarse|ls_peak|running_mean|calc_rois'
python bench.py --steps $steps --warmup 3 > $o/bench_c2_$tag.json 2> $o/bench_c2_$tag.err || { echo "bench failed"; tail -5 $o/bench_c2_$tag.err; exit 1; }
cat $o/bench_c2_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > $o/bench_c2_${tag}_ref.json 2>> $o/bench_c2_$tag.err
python bench.py --steps 3 --warmup 3 --no-cpu > $o/plain_$tag.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 400 --csv --log-file $o/launches_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > $o/ncu_ll_$tag.log 2>&1
# ring prefill = 10 step_signals x 5 kernels; each warm-up step = 6 kernels -> skip 50 + 6, capture one full step
ncu --set full --clock-control none --import-source on -k "$KF" --launch-skip 56 --launch-count 6 -f -o $o/c2_$tag \
    python bench.py --steps 1 --warmup 3 --no-cpu > $o/ncu_full_$tag.log 2>&1
tail -3 $o/ncu_full_$tag.log | cut -c1-300
