#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few numbers we judge by."""
import csv, subprocess, sys, io, json
KEYS = ['gpu__time_duration.sum','sm__cycles_elapsed.max','dram__bytes_read.sum','dram__bytes_write.sum',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
 'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_bytes.sum',
 'l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_xbar2l1tex_read_sectors.sum.pct_of_peak_sustained_elapsed',
 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size',
 'smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu.sum','smsp__cycles_active.avg',
 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum','l1tex__f_wavefronts.sum','lts__t_sectors.sum.pct_of_peak_sustained_elapsed']
def main(path):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for d in data:
        r = {'kernel': d[hdr.index('Kernel Name')]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k); r[k] = d[i] + ' ' + units[i]
        st = {}
        for i,h in enumerate(hdr):
            if 'average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
                try: v=float(d[i])
                except: continue
                if v >= 0.5: st[h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')] = round(v,2)
        r['stalls_per_issue'] = dict(sorted(st.items(), key=lambda kv:-kv[1]))
        res.append(r)
    print(json.dumps(res, indent=1))
if __name__ == '__main__':
    main(sys.argv[1])
