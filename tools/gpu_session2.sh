set -u
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q -k "dec or zstd or frames" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log ); tail -4 $O/pytest_gpu.log
timeout 300 python tools/enc_times.py 1 1 --decode 2>&1 | tee $O/enc_times_L1.log
B2C_DEC_SEQ=l2 timeout 300 python tools/enc_times.py 1 1 --decode 2>&1 | tee $O/enc_times_L1_seql2.log
