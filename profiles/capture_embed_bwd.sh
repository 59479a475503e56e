set -e
timeout 500 ncu --set full --clock-control none --import-source on -k regex:embed_bwd_fused -s 3 -c 1 -o gpurun_out/embed_bwd_v2 python profiles/ab_embed.py > gpurun_out/ncu_embed.log 2>&1 || tail -5 gpurun_out/ncu_embed.log
ls -la gpurun_out/*.ncu-rep
