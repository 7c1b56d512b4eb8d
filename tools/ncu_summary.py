#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of counters DESIGN.md quotes."""
import csv, subprocess, sys
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum', 'lts__t_sectors.sum',
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
        try:        # achieved bandwidths against the B200 peaks (HBM: MEASURED_PEAKS.json copy bandwidth; L2: dr_probe_l2_read, bench.py)
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            num = lambda k: float(r[idx[k]].replace(",", ""))
            dur_s = num("gpu__time_duration.sum") * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[units[idx["gpu__time_duration.sum"]]]
            dram = num("dram__bytes_read.sum") * scale[units[idx["dram__bytes_read.sum"]]] + num("dram__bytes_write.sum") * scale[units[idx["dram__bytes_write.sum"]]]
            l2 = 32.0 * num("lts__t_sectors.sum")
            print(f"{'achieved HBM GB/s (dram bytes / duration) vs 6552.6 measured peak':72s} {dram / dur_s / 1e9:.1f} GB/s = {100 * dram / dur_s / 6552.6e9:.2f} %")
            print(f"{'achieved L2 GB/s (32 B x lts__t_sectors / duration) vs 19000 measured peak':72s} {l2 / dur_s / 1e9:.1f} GB/s = {100 * l2 / dur_s / 19.0e12:.2f} %")
            wi = num("smsp__inst_executed.sum")
            clk = num("sm__cycles_elapsed.avg.per_second") * {"Ghz": 1e9, "Mhz": 1e6, "hz": 1.0}.get(units[idx["sm__cycles_elapsed.avg.per_second"]], 1e9)
            print(f"{'issue roof: warp instructions / (148 SMs x 4 schedulers x clock x duration)':72s} {wi / (148 * 4 * clk * dur_s):.3f}")
        except (KeyError, ValueError):
            pass
        st = [(float(r[i].replace(',', '')) if r[i] not in ('', 'n/a') else 0.0, h) for h, i in idx.items()
              if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio')]
        st.sort(reverse=True)
        print("top warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active):")
        for v, h in st[:8]:
            print(f"    {h[len('smsp__average_warps_issue_stalled_'):]:55s} {v:.3f}")
        print()
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
