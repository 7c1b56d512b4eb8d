#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of counters DESIGN.md quotes."""
import csv, subprocess, sys
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum', 'sm__cycles_elapsed.avg.per_second']
def main(rep, title=""):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(title)
    for r in rows[2:]:
        for k in KEEP:
            if k in idx:
                print(f"{k:72s} {r[idx[k]]} {units[idx[k]]}")
        st = [(float(r[i].replace(',', '')) if r[i] not in ('', 'n/a') else 0.0, h) for h, i in idx.items()
              if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio')]
        st.sort(reverse=True)
        print("top warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active):")
        for v, h in st[:8]:
            print(f"    {h[len('smsp__average_warps_issue_stalled_'):]:55s} {v:.3f}")
        print()
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
