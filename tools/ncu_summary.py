#!/usr/bin/env python
"""Prints the judged subset of an `ncu --page raw --csv` dump (one block per profiled launch)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'sm__cycles_elapsed.max',
        'lts__t_sectors.sum', 'lts__t_sectors.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sectors.sum', 'l1tex__t_requests.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.avg.per_cycle_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
STALL = 'smsp__average_warps_issue_stalled_'


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("== %s  (id %s)" % (d.get('Kernel Name', '?')[:100], d.get('ID', '?')))
        for w in WANT:
            if w in d and d[w] != '':
                print("  %-78s %16s %s" % (w, d[w], units[hdr.index(w)]))
        st = [(k, d[k]) for k in hdr if k.startswith(STALL) and k.endswith('_per_issue_active.ratio') and d[k] not in ('', '0')]
        st.sort(key=lambda kv: -float(kv[1].replace(',', '')))
        for k, v in st[:8]:
            print("  stall %-72s %16s" % (k[len(STALL):-len('_per_issue_active.ratio')], v))


if __name__ == '__main__':
    main(sys.argv[1])
