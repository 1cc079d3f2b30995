set -u
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log ); tail -4 $O/pytest_gpu.log
timeout 300 python tools/enc_times.py 1 1 --frames --frames-decode --frame-bytes=4194304 2>&1 | tee $O/enc_times_L1_4m.log
timeout 300 python tools/enc_times.py 1 1 --frames --frames-decode 2>&1 | tee $O/enc_times_L1_1m.log
