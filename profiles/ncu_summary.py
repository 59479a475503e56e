"""Summarise an .ncu-rep (ncu --set full) into JSON lines: one object per kernel launch with the metrics that matter
for this path.  Usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.jsonl"""
import csv, json, subprocess, sys
KEYS = {
 'gpu__time_duration.sum': 'time', 'dram__bytes_read.sum': 'dram_read', 'dram__bytes_write.sum': 'dram_write',
 'lts__t_bytes.sum': 'l2_bytes', 'l1tex__t_bytes.sum': 'l1_bytes', 'launch__registers_per_thread': 'regs',
 'sm__warps_active.avg.pct_of_peak_sustained_active': 'occupancy_pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_pct',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed': 'l1_pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2_pct',
 'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct', 'lts__t_sector_hit_rate.pct': 'l2_hit_pct',
 'l1tex__t_sector_hit_rate.pct': 'l1_hit_pct', 'lts__t_sectors_srcunit_tex_op_red.sum': 'l2_red_sectors',
 'lts__t_sectors_srcunit_tex_op_atom.sum': 'l2_atom_sectors',
 'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed': 'l2_atomic_unit_pct',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'stall_long_scoreboard',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio': 'stall_lg_throttle',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'stall_barrier',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio': 'stall_wait',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'stall_short_scoreboard',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'stall_math_throttle',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio': 'stall_not_selected',
 'smsp__inst_executed.sum': 'warp_insts', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum': 'ld_requests',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum': 'ld_sectors', 'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum': 'red_requests',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum': 'red_sectors', 'launch__grid_size': 'grid', 'launch__block_size': 'block',
}
TRAFFIC = '--traffic' in sys.argv
if TRAFFIC:
    sys.argv.remove('--traffic')
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
traffic = {}
for r in rows[2:]:
    d = {'kernel': r[h.index('Kernel Name')].split('(')[0][:60]}
    for k, short in KEYS.items():
        if k in h:
            i = h.index(k)
            try:
                d[short] = float(r[i].replace(',', ''))
            except ValueError:
                d[short] = r[i]
            if units[i] and short in ('time', 'dram_read', 'dram_write', 'l2_bytes', 'l1_bytes'):
                d[short + '_unit'] = units[i]
    if TRAFFIC:
        name = d['kernel'].replace('void ', '').replace('afb::', '').strip()
        tot = d.get('dram_read', 0.0) * UNIT.get(d.get('dram_read_unit', 'byte'), 1.0) + d.get('dram_write', 0.0) * UNIT.get(d.get('dram_write_unit', 'byte'), 1.0)
        traffic.setdefault(name, []).append(tot)
    else:
        print(json.dumps(d))
if TRAFFIC:
    print(json.dumps({'source': sys.argv[1], 'what': 'dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over captured launches)',
                      'kernels': {k: sum(v) / len(v) for k, v in traffic.items()}}, indent=1))
