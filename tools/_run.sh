for ms in 1024 1280; do
  LFB_MS=$ms python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($ms, d['value'], d['ms_per_step']); [print('  %-36s %.4f'%kv) for kv in d['roofline']['kernel_ms_serial'].items() if 'flux' in kv[0]]"
done
