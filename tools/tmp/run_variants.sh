set -x
for v in cur d1 d2 fuse1 fuse2 fuse4 cur fuse2; do
  B2C_LIB=$PWD/variants/$v.so timeout 300 python tools/kernel_times.py 2>&1 | tail -4
done
B2C_LIB=$PWD/variants/fuse2.so timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_zstd_gpu.py::test_native_library_loaded 2>&1 | tail -3
