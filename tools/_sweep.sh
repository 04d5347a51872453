(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
python bench.py --skip-cpu --skip-ntt --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), {k:round(v,3) for k,v in d['roofline']['phases_ms'].items()})"
