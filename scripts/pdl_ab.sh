run() { python bench.py --steps 200 --warmup 10 --no-cpu --skip-e2e 2>gpurun_out/pdl_err_$1.log | grep '^{' | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t); print('$1', 'ms/round', round(d['ms_per_step']*1e3,2), {k:d[k] for k in d if k in ('k1_ms','k2_ms','value')})
except Exception as e: print('$1', 'no line', repr(t[:200]))"; tail -2 gpurun_out/pdl_err_$1.log; }
run dflt; KTN_PDL=0 run nopdl; KTN_K1_EVENT_EVERY=1 run every1; run dflt2; KTN_K1_EVENT_EVERY=64 run every64
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
