"""Top stall locations of one kernel from an .ncu-rep source page:  python tools/ncu_hot.py rep.ncu-rep <launch-index> [N]"""
import csv, subprocess, sys


def main(path, which=0, top=40):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            secs.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    sec = secs[which]
    hdr = sec['rows'][0]
    data = [r for r in sec['rows'][1:] if len(r) >= len(hdr)]
    si, so, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    stalls = [(h, hdr.index(h)) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[si]) for r in data)
    print(sec['name'][:80], '| samples', tot, '| instructions', len(data))
    agg = {}
    for r in data:
        for h, i in stalls:
            v = int(r[i])
            if v:
                agg[h] = agg.get(h, 0) + v
    print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:top]
    for i in sorted(idx):
        r = data[i]
        st = sorted(((h, int(r[j])) for h, j in stalls if int(r[j]) > 0), key=lambda kv: -kv[1])[:2]
        print(i, r[si], r[ie], r[so].strip()[:70], st)


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
