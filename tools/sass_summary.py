"""Per-kernel counts of the SASS mnemonics that prove the Blackwell path (tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM,
tcgen05.commit = UTCBAR, cp.async.bulk = UBLKCP, mbarrier = SYNCS, redux = REDUX) in the shipped library.
usage: python tools/sass_summary.py [libmuzero_b200.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'muzero_b200', 'libmuzero_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
OPS = ['UTCHMMA', 'LDTM', 'UTCBAR', 'UBLKCP', 'UTMALDG', 'SYNCS', 'REDUX', 'HMMA', 'DFMA', 'ELECT']
cur, counts, arch = None, collections.OrderedDict(), set()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r'arch = (sm_\w+)', line)
    if m:
        arch.add(m.group(1))
    if cur:
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1).split('.')[0]
            counts[cur]['_total'] += 1
            if op in OPS:
                counts[cur][op] += 1
print(f'# {os.path.basename(lib)}: cuobjdump -sass, architectures {sorted(arch)}')
print(f'{"kernel":78s} {"instr":>7s} ' + ' '.join(f'{o:>8s}' for o in OPS))
tot = collections.Counter()
for k, c in counts.items():
    name = demangle(k)
    name = re.sub(r'\((?:mz::|const |int|float|unsigned|double|PoolDev|ConvParams|MlpTcParams|HeadsParams)[^()]*(\([^()]*\)[^()]*)*\)$', '', name)
    name = name.replace('(int)', '').replace('(bool)', '')[:78]
    print(f'{name:78s} {c["_total"]:7d} ' + ' '.join(f'{c[o]:8d}' for o in OPS))
    tot.update(c)
print(f'{"TOTAL":78s} {tot["_total"]:7d} ' + ' '.join(f'{tot[o]:8d}' for o in OPS))
