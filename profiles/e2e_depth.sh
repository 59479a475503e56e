#!/bin/bash
# e2e leg of bench.py vs the number of batches in flight (device buffer sets of HostInputPipeline)
for dp in $1; do
timeout 300 python bench.py --no-variants --no-weak --no-cpu-baseline --e2e-depth $dp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; p=e['host_label_packing'] or {}
print(json.dumps({'depth': $dp, 'e2e_slices_per_s': round(e['value']), 'ms_per_step': round(e['ms_per_step'],2), 'h2d_gbs': round(e['h2d_gbs_per_rank'],1), 'pack_ms': round(p.get('pack_ms_per_batch',0),2), 'worker_ms': round(p.get('worker_ms_per_batch',0),2), 'value': round(d['value'])}))"
done
