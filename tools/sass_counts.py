"""Per-kernel SASS instruction counts of the tcgen05 / TMA kernels (cuobjdump -sass of the built objects): the mnemonics
that prove the tensor-core path (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, RED = fp32 reductions, LDGSTS = cp.async).  usage: python tools/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, 'sr-gan_b200', 'csrc')
KEYS = ('UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'UTCBAR', 'SYNCS', 'LDGSTS', 'RED', 'ATOMG', 'STG', 'LDG', 'HSETP2', 'HSET2', 'FFMA', 'HFMA2', 'MUFU')
print('# cuobjdump -sass of sr-gan_b200/csrc/*.o (nvcc 12.9, sm_100a); static instruction counts per kernel')
for obj in ('umma_conv.o', 'graph_ops.o', 'skinny.o', 'elementwise.o', 'coef_step.o', 'simt_conv.o'):
    out = subprocess.run(['cuobjdump', '-sass', os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    cur, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'\(.*', '', cur.replace('(anonymous namespace)::', '')).replace('void ', '')
            counts[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', line)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    print(f'\n== {obj}')
    for k, c in counts.items():
        if obj != 'umma_conv.o' and total[k] < 400 and not any(c[x] for x in ('UTCHMMA', 'UTMALDG')):
            continue
        print(f'{total[k]:6d} instr  ' + ' '.join(f'{x}={c[x]}' for x in KEYS if c[x]) + f'   {k[:90]}')
