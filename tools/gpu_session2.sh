set -u
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log ); tail -5 $O/pytest_gpu.log
for L in 1 2 3; do timeout 300 python tools/enc_times.py $L 1 --frames 2>&1 | tee $O/enc_times_frames_L$L.log; done
B2C_ENC_SIDE=0 timeout 300 python tools/enc_times.py 1 1 2>&1 | tee $O/enc_times_L1_noside.log
