"""SASS mnemonic counts per kernel family of the built library -> profiles/r2_sass_counts.txt (evidence that the hot
kernels are tcgen05 / TMEM / bulk-copy code: UTCHMMA, LDTM, UBLKCP, UTCBAR).  Runs on the CPU box (cuobjdump only).

    python tools/sass_counts.py [output file]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'dlwp_cs_b200', 'lib', 'libdlwpcs.so')
PAT = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTCBAR', 'UTCATOMSWS', 'LDGSTS', 'UTMALDG', 'UTMASTG', 'HMMA', 'SYNCS', 'UTCCP']
sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        mm = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if mm:
            op = mm.group(1).split('.')[0]
            counts[cur]['_n'] += 1
            if op in PAT:
                counts[cur][op] += 1
names = list(counts)
dem = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
fam = collections.OrderedDict()
for f, d in zip(names, dem):
    d = d.replace('(anonymous namespace)::', '').replace('dlwpcs::', '').replace('void ', '')
    key = re.sub(r'<.*>', '<...>', re.sub(r'\(.*', '', d))
    e = fam.setdefault(key, [0, collections.Counter()])
    e[0] += 1
    e[1].update(counts[f])
tot = collections.Counter()
lines = ['SASS instruction counts of dlwp_cs_b200/lib/libdlwpcs.so (cuobjdump -sass, sm_100a cubins), per kernel family',
         '(instantiations summed; tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, cp.async.bulk = UBLKCP, tcgen05.commit / mbarrier = UTCBAR / SYNCS,',
         'cp.async = LDGSTS, tcgen05.alloc = UTCATOMSWS; tensor-map TMA would be UTMALDG / UTMASTG, legacy mma.sync would be HMMA)', '',
         '%-30s %5s %9s  %s' % ('kernel family', 'inst.', 'instr.', 'mnemonics')]
for k, (n, c) in fam.items():
    tot.update(c)
    lines.append('%-30s %5d %9d  %s' % (k[:30], n, c['_n'], ' '.join('%s=%d' % (p, c[p]) for p in PAT if c[p])))
lines += ['', 'total: ' + ' '.join('%s=%d' % (p, tot[p]) for p in PAT)]
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r2_sass_counts.txt')
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
