for l in 2 1; do
LFB_LANES=$l python bench.py --steps 50 --warmup 3 --no-cpu --no-gp 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lanes', $l, 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'emcee', round(d['emcee_steps_per_s'],1), 'dev', round(d['emcee']['device_resident_steps_per_s'],1))"
done
