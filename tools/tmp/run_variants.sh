set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for v in base v1 v2 v4 v5 v3 base; do
  B2C_LIB=$PWD/variants/$v.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
done
for ov in 0 1 2; do
  echo "== overlap $ov"
  B2C_OVERLAP=$ov B2C_LIB=$PWD/variants/ov.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
done
echo "== free-running tables, overlap 0 / 2"
B2C_OVERLAP=0 B2C_LIB=$PWD/variants/fr.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
B2C_OVERLAP=2 B2C_LIB=$PWD/variants/fr.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
echo "== overlap 2, side grid 8/SM; same + free-running + 8 sub-batches"
B2C_OVERLAP=2 B2C_LIB=$PWD/variants/ovg8.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
B2C_OVERLAP=2 B2C_LIB=$PWD/variants/ovg8fr8.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
echo "== overlap 2, 8 sub-batches"
B2C_OVERLAP=2 B2C_LIB=$PWD/variants/ov8.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
B2C_OVERLAP=2 B2C_LIB=$PWD/variants/ov.so timeout 600 python -m pytest tests/test_zstd_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m pytest tests/test_zstd_gpu.py tests/test_huf0_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name 'regex:b2c_zstd_(tables|pack)_kernel' -c 2 -f -o gpurun_out/prof_r01h_tables_pack python tools/kernel_times.py 16384 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
