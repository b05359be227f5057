"""Summarise .ncu-rep captures (read here, no GPU) into a small text file for profiles/."""
import csv, subprocess, sys, io

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']


def summarise(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = []
    seen = set()
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')]
        if name in seen:                      # one block per kernel: the first captured launch
            continue
        seen.add(name)
        lines.append(f"kernel: {name}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"  {w:72s} {vals[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
                try:
                    stalls.append((float(vals[i]), h))
                except ValueError:
                    pass
        for v, h in sorted(stalls, reverse=True)[:5]:
            lines.append(f"  stall {h:66s} {v:18.2f} %")
    return "\n".join(lines)


if __name__ == '__main__':
    for p in sys.argv[1:]:
        print(f"# {p}")
        print(summarise(p))
        print()
