"""SASS opcode histogram of libmmnas_b200.so (no GPU needed): per kernel, how many tcgen05 / TMEM / TMA / legacy-MMA
instructions the shipped binary contains.  The PTX names never appear in SASS (B200_PROFILING.md):
   tcgen05.mma -> UTC*MMA   tcgen05.ld / st -> LDTM / STTM   TMA -> UTMALDG / UTMASTG / UTMAREDG / UTMAPF   mma.sync -> HMMA
    python scripts/sass_histogram.py [out.json]"""
import collections, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'mmnas_b200', 'lib', 'libmmnas_b200.so')
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTMAPF', 'UBLKCP', 'HMMA', 'SYNCS', 'UCGABAR_ARV',
         'UCGABAR_WAIT', 'MUFU', 'FFMA', 'FFMA2', 'LDG', 'STG', 'LDS', 'STS', 'ATOMG', 'REDG', 'RED']
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(anonymous namespace\)::', '', name).split('(')[0].replace('void ', '')
        cur = kernels.setdefault(name, collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
    if m and cur is not None:
        op = m.group(1)
        cur['_total'] += 1
        if op in WATCH:
            cur[op] += 1
res = {k: dict(v) for k, v in kernels.items()}
for k, v in res.items():
    tc = ' '.join('%s=%d' % (o, v[o]) for o in WATCH[:11] if v.get(o))
    print('%-70s %6d instr  %s' % (k[:70], v['_total'], tc))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], 'w'), indent=1)
