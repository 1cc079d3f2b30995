#!/bin/bash
# One GPU-box session: tests, per-kernel times, bench lines, launch list, ncu captures. Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log )
tail -3 $O/pytest_gpu.log
for L in 1 2 3; do
  EXTRA="--frames"; [ $L = 1 ] && EXTRA="--decode --s2 --s2stream --frames --frames-decode"
  timeout 300 python tools/enc_times.py $L 1 $EXTRA > $O/enc_times_L$L.log 2>&1
  cat $O/enc_times_L$L.log
done
timeout 300 python tools/huf_times.py > $O/huf_times.log 2>&1; cat $O/huf_times.log
SECONDS=0
timeout 900 python bench.py > $O/bench_L1.json 2> $O/bench_L1.err; tail -1 $O/bench_L1.json | cut -c1-600; echo "bench default took ${SECONDS}s"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -1 $O/bench_ref.json | cut -c1-300
timeout 900 python bench.py --level 2 --no-cpu-baseline > $O/bench_L2.json 2> $O/bench_L2.err; tail -1 $O/bench_L2.json | cut -c1-400
timeout 600 ncu --kernel-name regex:b2c_ --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2c_lz_parse1 -s 2 -c 1 -f -o $O/prof_parse1 \
   python tools/enc_times.py 1 0.5 > $O/ncu_parse1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2c_zstd_pack -s 2 -c 1 -f -o $O/prof_pack \
   python tools/enc_times.py 1 0.5 > $O/ncu_pack.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:b2c_zstd_dec_(seq|lit|exec)" -s 6 -c 3 -f -o $O/prof_dec \
   python tools/enc_times.py 1 0.5 --decode > $O/ncu_dec.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2c_lz_parse2 -s 2 -c 1 -f -o $O/prof_parse2 \
   python tools/enc_times.py 2 0.5 > $O/ncu_parse2.log 2>&1
ls -la $O | head -40
