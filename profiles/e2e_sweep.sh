#!/bin/bash
# e2e leg of bench.py for several upload group sizes / packing chunk sizes: one line each
# usage: bash profiles/e2e_sweep.sh "8 16" "16 32 64"      (upload groups, volumes per packing call)
for g in $1; do for pv in $2; do
AFB_PACK_VOLUMES=$pv timeout 300 python bench.py --no-variants --no-weak --e2e-group $g 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; p=e['host_label_packing']
print(json.dumps({'group': $g, 'pack_volumes': $pv, 'threads': p['threads'], 'e2e_slices_per_s': round(e['value']), 'ms_per_step': round(e['ms_per_step'],2), 'h2d_gbs': round(e['h2d_gbs_per_rank'],1), 'pack_ms': round(p['pack_ms_per_batch'],2), 'worker_ms': round(p['worker_ms_per_batch'],2), 'value': round(d['value'])}))"
done; done
