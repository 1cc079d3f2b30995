set -u
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log ); tail -4 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_L1.json 2> $O/bench_L1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_L1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e_chunk_apis'])
PY
