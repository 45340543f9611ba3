"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a compact per-launch table."""
import csv, subprocess, sys
KEYS = [('Grid Size', 'grid'), ('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrMB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%act'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor%'),
        ('sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'hmma%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__occupancy_limit_shared_mem', 'occ_smem'), ('launch__registers_per_thread', 'regs'),
        ('lts__t_sector_hit_rate.pct', 'l2hit%'), ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'bankconf'),
        ('launch__shared_mem_per_block_dynamic', 'smemKB')]


def main(path, filt=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), n, units[hdr.index(k)]) for k, n in KEYS if k in hdr]
    print(' | '.join(['kernel'] + ['%s[%s]' % (n, u) for _, n, u in cols]))
    ki = hdr.index('Kernel Name')
    for r in rows[2:]:
        if filt and filt not in r[ki]:
            continue
        name = r[ki].split('(')[0].split('::')[-1][:24]
        print(' | '.join([name] + [r[i] for i, _, _ in cols]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
