for pad in 0 10 45; do
  LFB_FLUX_PAD_KB=$pad python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($pad, d['value'], d['ms_per_step'], d['roofline']['kernel_ms_serial']['flux_kernel'])"
done
