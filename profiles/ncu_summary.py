"""Summarise an `ncu --page raw --csv` export: one block per profiled launch with the metrics the roofline uses."""
import csv
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_lsu.sum', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'launch__occupancy_limit_blocks']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('-----')
        for w in WANT:
            if w in idx:
                print(f"  {w:62s} {r[idx[w]]} {units[idx[w]]}")
        try:
            t = float(r[idx['gpu__time_duration.sum']].replace(',', ''))
            u = units[idx['gpu__time_duration.sum']]
            t_s = t * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}.get(u, 1e-9)
            rd = float(r[idx['dram__bytes_read.sum']].replace(',', ''))
            wr = float(r[idx['dram__bytes_write.sum']].replace(',', ''))
            scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            rd *= scale.get(units[idx['dram__bytes_read.sum']], 1)
            wr *= scale.get(units[idx['dram__bytes_write.sum']], 1)
            print(f"  => dram traffic {(rd + wr) / 1e6:.1f} MB (read {rd / 1e6:.1f}, write {wr / 1e6:.1f}); {(rd + wr) / t_s / 1e9:.0f} GB/s under ncu")
        except Exception as e:  # noqa: BLE001
            print("  (no traffic summary:", e, ")")


if __name__ == '__main__':
    main(sys.argv[1])
