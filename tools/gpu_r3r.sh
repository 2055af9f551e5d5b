#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 --no-ref-gpu --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['roofline']['families_achieved']
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v.get('GB/s',0)) for k,v in f.items()})"
true
