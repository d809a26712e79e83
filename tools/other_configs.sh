#!/bin/bash
# window-pipeline timings on the shapes of BASELINE configs 1/5, 3 and 4 (signals-only; bench.py measures config 2)
o=gpurun_out/other_configs_${1:-r1k}.txt
: > $o
echo "== config 5 shape per GPU: 8192 streams (65536 / 8), W=300, BUTTER + LS (F = n), xcorr, one window per stream per step" >> $o
python tools/bench_window.py --S 8192 --T 1 --W 300 --methods FILTER_BUTTER --transform PGRAM_LS --windows last 2>&1 | tail -5 >> $o
echo "== config 5, all 65536 streams on ONE GPU" >> $o
python tools/bench_window.py --S 65536 --T 1 --W 300 --methods FILTER_BUTTER --transform PGRAM_LS --windows last --iters 5 2>&1 | tail -5 >> $o
echo "== config 3: 4096 streams, irregular timestamps, LS on a 2048-frequency grid, no interpolation" >> $o
python tools/bench_window.py --S 4096 --T 1 --W 300 --methods "" --transform PGRAM_LS --ls-num-freqs 2048 --irregular --windows last 2>&1 | tail -5 >> $o
echo "== config 4: 1024 streams, 120 fps, W=1200, INTERP_CUBIC + FILTER_BUTTER, LS, xcorr over 2399 lags" >> $o
python tools/bench_window.py --S 1024 --T 1 --W 1200 --fps 120 --methods INTERP_CUBIC,FILTER_BUTTER --transform PGRAM_LS --irregular --windows last 2>&1 | tail -5 >> $o
echo "== config 4 with DFT_RFFT" >> $o
python tools/bench_window.py --S 1024 --T 1 --W 1200 --fps 120 --methods INTERP_CUBIC,FILTER_BUTTER --transform DFT_RFFT --irregular --windows last 2>&1 | tail -5 >> $o
echo "== config 1 shape batched: 256 streams x 32 frames, W=300, BUTTER (0.7-4 Hz) + LS, every frame" >> $o
python tools/bench_window.py --S 256 --T 32 --W 300 --methods FILTER_BUTTER --transform PGRAM_LS 2>&1 | tail -5 >> $o
cat $o
