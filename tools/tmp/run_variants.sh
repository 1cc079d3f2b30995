set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for v in base v1 v2 v4 v5 v6 v3 v7 base v1; do
  B2C_LIB=$PWD/variants/$v.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
done
for ov in 0 1 2; do
  echo "== overlap $ov"
  B2C_OVERLAP=$ov B2C_LIB=$PWD/variants/ov.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
done
echo "== free-running tables, overlap 0 / 2"
B2C_OVERLAP=0 B2C_LIB=$PWD/variants/fr.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
B2C_OVERLAP=2 B2C_LIB=$PWD/variants/fr.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
echo "== overlap 2, 8 sub-batches"
B2C_OVERLAP=2 B2C_LIB=$PWD/variants/ov8.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
B2C_OVERLAP=2 B2C_LIB=$PWD/variants/ov.so timeout 600 python -m pytest tests/test_zstd_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m pytest tests/test_zstd_gpu.py tests/test_huf0_gpu.py -x -q -m gpu 2>&1 | tail -3
