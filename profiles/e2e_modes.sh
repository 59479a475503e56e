#!/bin/bash
# e2e leg of bench.py for the upload modes (0 int64 upload, 1 all labels packed on the host, 2 split upload) at a given number of
# packing threads (emulates hosts with few cores per GPU).   usage: bash profiles/e2e_modes.sh "<threads...>" "<modes...>"
for nt in $1; do for m in $2; do
AFB_NARROW_THREADS=$nt timeout 300 python bench.py --no-variants --no-weak --no-cpu-baseline --e2e-narrow $m 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; p=e['host_label_packing'] or {}
print(json.dumps({'threads': $nt, 'mode': $m, 'e2e_slices_per_s': round(e['value']), 'ms_per_step': round(e['ms_per_step'],2), 'h2d_gbs': round(e['h2d_gbs_per_rank'],1), 'h2d_bytes': e['h2d_bytes_per_step'], 'pack_ms': round(p.get('pack_ms_per_batch',0),2), 'split': e.get('split_upload')}))"
done; done
