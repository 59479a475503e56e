#!/bin/bash
# Final round-2 refresh with the last library: ncu launch list of the eager step + full capture of the embedding / one-hot kernels.
out=gpurun_out
mkdir -p $out
if [ "$1" != "embed-only" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r2_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --graph off --no-cpu-baseline --no-breakdown --no-variants --e2e-steps 2 > $out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
fi
ncu --set full --clock-control none -k 'regex:(embed_|onehot_)' --launch-count 12 -f -o $out/r2_prof_embed_onehot python profiles/capture_embed_onehot.py > $out/r2_ncu_embed.log 2>&1; echo "ncu embed rc=$?"
ls -la $out/*.ncu-rep
