set -u
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log ); tail -4 $O/pytest_gpu.log
timeout 300 python tools/huf_times.py 2>&1 | tee $O/huf_times.log
B2C_DEC=onewarp timeout 300 python tools/huf_times.py 2>&1 | tee $O/huf_times_onewarp.log
timeout 300 python tools/enc_times.py 1 1 --s2stream 2>&1 | tee $O/enc_times_L1.log
