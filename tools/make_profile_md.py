"""Regenerate the tables of profiles/r<N>_tc_kernels.md and profiles/r<N>_traffic.json from the ncu artefacts in gpurun_out/
(r<N>_tc*.ncu-rep, [r<N>_wgrad.ncu-rep,] r<N>_launches.csv, r<N>_train_launches.csv).  Prints the tables as markdown.

    python tools/make_profile_md.py [round, default 2]"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'gpurun_out')
K = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
     'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
     'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'launch__shared_mem_per_block_dynamic',
     'launch__registers_per_thread']
NAMES = ['conv_2d_1', 'conv_2d_1_2', 'conv_2d_2', 'conv_2d_2_2', 'conv_2d_5_2', 'conv_2d_5', 'conv_2d_6_2', 'conv_2d_6',
         'conv_2d_7', 'conv_2d_7_2', 'conv_2d_8']


def table(path, names):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h, rs = r[0], r[2:]
    idx = [h.index(k) for k in K]
    ki = h.index('Kernel Name')
    lines = ['| layer | kernel | us | DRAM rd MB | DRAM wr MB | DRAM % | tensor pipe % (elapsed) | SM % | L2 hit % | smem KB | regs |',
             '|---|---|---|---|---|---|---|---|---|---|---|']
    traffic = {}
    if len(rs) == len(names) - 1:        # the 1x1 output layer runs inside the launch of the layer in front of it
        names = names[:-2] + [names[-2] + '+' + names[-1]]
    for n, row in zip(names, rs):
        v = [float(row[i]) for i in idx]
        kn = row[ki].split('(')[0].split('::')[-1]
        lines.append('| %s | `%s` | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.0f | %d |'
                     % (n, kn, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]))
        traffic[n] = (v[1] + v[2]) * 1e6
    return '\n'.join(lines), traffic


def shares(path, top=14):
    agg, hdr = collections.OrderedDict(), None
    for r in csv.reader(open(path)):
        if hdr is None:
            if 'Kernel Name' in r:
                hdr = r
            continue
        if len(r) < len(hdr) or r[hdr.index('Metric Name')] != 'gpu__time_duration.sum':
            continue
        v, u = float(r[hdr.index('Metric Value')]), r[hdr.index('Metric Unit')]
        v = v / 1000.0 if u in ('ns', 'nsecond') else v
        a = agg.setdefault(r[hdr.index('Kernel Name')].split('(')[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = ['| kernel | launches | total us | share |', '|---|---|---|---|']
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        lines.append('| `%s` | %d | %.1f | %.1f %% |' % (k[:100], a[0], a[1], 100 * a[1] / tot))
    return '\n'.join(lines)


if __name__ == '__main__':
    rnd = sys.argv[1] if len(sys.argv) > 1 else '2'
    rep = os.path.join(G, 'r%s_tc.ncu-rep' % rnd)
    if not os.path.exists(rep):
        rep = os.path.join(G, 'r%s_tc_final.ncu-rep' % rnd)
    t1, tr = table(rep, NAMES)
    json.dump({'source': 'profiles/r%s_tc_kernels.md (ncu --set full, batch 64, C48, bf16)' % rnd, 'batch': 64,
               'dram_bytes_per_launch': tr}, open(os.path.join(ROOT, 'profiles', 'r%s_traffic.json' % rnd), 'w'), indent=1)
    for f in ('r%s_launches.csv' % rnd, 'r%s_train_launches.csv' % rnd):
        shutil.copy(os.path.join(G, f), os.path.join(ROOT, 'profiles', f))
    print('## rollout launch list\n' + shares(os.path.join(G, 'r%s_launches.csv' % rnd)))
    print('\n## forward kernel\n' + t1)
    print('\n## training launch list\n' + shares(os.path.join(G, 'r%s_train_launches.csv' % rnd), 26))
    wg = os.path.join(G, 'r%s_wgrad.ncu-rep' % rnd)
    if os.path.exists(wg):
        print('\n## wgrad kernel\n' + table(wg, list(reversed(NAMES)))[0])
