python -m pytest tests -q -m gpu -x 2>&1 | tail -30
python bench.py --nx 128 --ny 128 --nz 32 --ppc 64 --steps 3 --warmup 2 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e9,'G/s', d['ms_per_step'],'ms', d['roofline']['phases_ms'], d['roofline']['frac'], d['checks'])"
