set -u
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log ); tail -4 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench_L1.json 2> $O/bench_L1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_L1.json').read().strip().splitlines()[-1])
print(d['value'], d['decode']['value'])
for k,v in d['s2'].items(): print(k, {a: round(b,1) if isinstance(b,float) else b for a,b in v.items() if a!='note'})
print(d['e2e_chunk_apis'])
PY
B2C_DEC=onewarp timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('onewarp:', {k: round(v['decode_gbs'],1) for k,v in d['s2'].items() if 'decode_gbs' in v})"
