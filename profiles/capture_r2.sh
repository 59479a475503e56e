#!/bin/bash
# Round-2 capture, run on the GPU box through gpurun:   gpurun --timeout 1200 -- 'bash profiles/capture_r2.sh'
#   1. bench line (the driver's command)          2. reference arm          3. ncu launch list of the eager step
#   4. one `ncu --set full` capture of one whole step at 16 volumes (short replays) + the embedding / one-hot kernels
# Numbers printed under ncu are never bench values; they are kept for the per-launch shares and counters only.
out=gpurun_out
mkdir -p $out
python bench.py > $out/r2_bench_final_n1.json 2> $out/r2_bench_final_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > $out/r2_bench_final_reference.json 2> $out/r2_bench_final_reference.err; echo "reference rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r2_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --graph off --no-cpu-baseline --no-breakdown --no-variants --e2e-steps 2 > $out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:^(min_grad|slice_|view_|volume_min)' --launch-skip 24 --launch-count 8 -f -o $out/r2_prof_step \
    python bench.py --volumes 16 --steps 2 --warmup 3 --graph off --no-cpu-baseline --no-breakdown --no-variants --e2e-steps 2 > $out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none -k 'regex:^(embed_|onehot_)' --launch-count 12 -f -o $out/r2_prof_embed_onehot python - > $out/r2_ncu_embed.log 2>&1 <<'PY'
import sys
sys.path.insert(0, ".")
import torch
import acquisition_focus_b200 as afb
from acquisition_focus_b200 import synthetic as cases
dev = torch.device("cuda", 0)
B, V = 2, 6
case0 = cases.embed_case(128, 16, V, B, seed=300)
gas = [a.to(dev).requires_grad_(True) for a in case0["affines"]]
cfgs = ((16, 128), (32, 64), (64, 32), (128, 16), (256, 8), (256, 4))
xs = [cases.randn((B, V * c, S, S), 500 + S).to(dev).requires_grad_(True) for c, S in cfgs]
outs = afb.embed_slices_multi(xs, torch.stack(gas, 0), V)                      # one launch: all six stages
torch.autograd.backward(outs, [torch.ones_like(o) for o in outs])              # one launch: all six stages
o0 = afb.embed_slices(xs[0], torch.stack(gas, 0), V)                            # stage 0 alone: zero kernel + slab kernel
case = cases.atm_case(128, 8, 6, seed=43)
params = torch.stack(case["params"], 1).to(dev).requires_grad_(True)
ys, yl, yi, ga, nii, th = afb.acquire_views_from_labels(case["lab"].to(dev).to(torch.uint8), case["image"].to(dev), case["nii"].to(dev),
                                                        torch.stack(case["gpre"], 1).to(dev), params,
                                                        torch.tensor([[1e-2, 0, 0, 0, 1e-2, 0, 0, 0, 0, 1.0]]).repeat(6, 1).to(dev), num_classes=8,
                                                        offset_clip=0.2, zoom_clip=0.0, spat=128, slice_fov_mm=[192.0, 192.0, 1.5],
                                                        slice_fov_vox=[128, 128, 1])
ys.sum().backward()
torch.cuda.synchronize()
PY
echo "ncu embed rc=$?"
