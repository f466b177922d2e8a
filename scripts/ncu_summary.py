"""Key metrics of every launch in an ncu report (raw page) as text + a json blob: python scripts/ncu_summary.py rep out_prefix"""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_op_local_ld.sum',
        'smsp__inst_executed_op_local_st.sum', 'sm__cycles_elapsed.max']
res = []
with open(out + '.txt', 'w') as f:
    for vals in rows[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        rec = {}
        for k in keys:
            if k in d:
                f.write('%-70s %s %s\n' % (k, d[k], u[k])); rec[k] = d[k] + ' ' + u[k]
        f.write('stalls (cycles per issued instruction):\n')
        for k in hdr:
            if 'smsp__average_warp' in k and 'issue_stalled' in k and 'not_issued' not in k:
                try:
                    v = float(d[k])
                    if v > 0.2:
                        nm = k.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '')
                        f.write('    %-50s %.2f\n' % (nm, v)); rec['stall_' + nm] = v
                except ValueError:
                    pass
        f.write('\n')
        res.append(rec)
json.dump(res, open(out + '.json', 'w'), indent=1)
print(open(out + '.txt').read()[:3000])
