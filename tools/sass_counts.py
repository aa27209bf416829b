"""Per-kernel counts of the tcgen05 / TMEM / TMA instructions in libagrl_b200.so (cuobjdump -sass; no GPU needed):
UTCHMMA / UTCQMMA (tcgen05.mma kind::f16 / kind::f8f6f4), LDTM (tcgen05.ld), UTMALDG (cp.async.bulk.tensor), UBLKCP
(cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), UTCATOMSWS / UTCCP misc.   python tools/sass_counts.py > profiles/r2/sass_counts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'agrl', 'pytorch_b200', 'libagrl_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', out)), capture_output=True, text=True).stdout.split('\n')
PAT = ['UTCHMMA', 'UTCQMMA', 'UTCOMMA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'UTMAPF', 'FFMA', 'HMMA']
rows, cur, k = [], None, -1
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        k += 1
        cur = collections.Counter()
        rows.append((names[k], cur))
        continue
    if cur is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        cur['_total'] += 1
        for p in PAT:
            if op.startswith(p):
                cur[p + ('.2CTA' if '.2CTA' in op else '')] += 1
print('libagrl_b200.so  %d bytes  %d kernels   (sm_100a SASS, cuobjdump %s)' % (os.path.getsize(lib), len(rows),
      subprocess.run(['cuobjdump', '--version'], capture_output=True, text=True).stdout.strip().split('\n')[-1]))
cols = sorted({c for _, r in rows for c in r if c != '_total'})
print('%-110s %7s ' % ('kernel', 'instr') + ' '.join('%12s' % c for c in cols))
for n, r in sorted(rows, key=lambda x: x[0]):
    short = re.sub(r'\(.*', '', n).replace('agrl::', '').replace('gemm::', '')
    print('%-110s %7d ' % (short[:110], r['_total']) + ' '.join('%12s' % (r[c] or '.') for c in cols))
