#!/bin/bash
# Round capture, run on the GPU box through gpurun:   gpurun --timeout 900 -- 'bash profiles/capture.sh v7'
#   1. GPU parity tests        2. bench line (the driver's command)      3. ncu launch list of the same command
#   4. one `ncu --set full` capture of one whole step (16 volumes x 6 views so that the replays stay short)
# Numbers printed under ncu are never bench values; they are kept for the per-launch shares and counters only.
tag=${1:-dev}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["clocks"])
for k, v in d["kernels"].items():
    print(f"  {v['ms']:.3f} ms {v['gbs']:8.0f} GB/s  {k}")
print(d["variants"])
PY
if [ "$2" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown --e2e-steps 1 > $out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:^(min_grad|slice_|view_|volume_min|onehot_|embed_)' --launch-skip 30 --launch-count 12 -f -o $out/prof_${tag}_step \
    python bench.py --volumes 16 --steps 2 --warmup 3 --no-cpu-baseline --no-breakdown --e2e-steps 1 > $out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
fi
if [ "$2" != "noncu" ] && [ -f bench_extra.py ]; then
python bench_extra.py > $out/extra_$tag.jsonl 2> $out/extra_$tag.err; echo "bench_extra rc=$?"; tail -12 $out/extra_$tag.jsonl | cut -c1-260
fi
