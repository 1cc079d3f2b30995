set -u
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log ); tail -5 $O/pytest_gpu.log
for L in 1 2 3; do E=""; [ $L = 1 ] && E="--s2"; timeout 300 python tools/enc_times.py $L 1 --frames $E 2>&1 | tee $O/enc_times_frames_L$L.log; done
