# SASS evidence (read on the CPU box): tcgen05 / TMA / TMEM mnemonics per kernel of libsrb200.so.
#   UTCHMMA = tcgen05.mma kind::f16, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, LDTM = tcgen05.ld,
#   UTMALDG / UTMASTG = cp.async.bulk.tensor load / store, SYNCS = mbarrier ops, ELECT = elect.sync
cuobjdump -sass sr-pytorch-lightning_b200/csrc/libsrb200.so 2>/dev/null | python3 -c "
import sys,re,collections,subprocess
cur=None
cnt=collections.defaultdict(collections.Counter)
pat=re.compile(r'\b(UTCHMMA|UTMALDG|UTMASTG|UTCBAR|UTCATOMSWS|LDTM|SYNCS|ELECT|HMMA|REDG)\b')
for l in sys.stdin:
    m=re.search(r'Function : (\S+)',l)
    if m: cur=m.group(1); continue
    if cur:
        for k in pat.findall(l): cnt[cur][k]+=1
for f,c in cnt.items():
    if any(k in c for k in ('UTCHMMA','UTMALDG','LDTM')):
        name=subprocess.run(['c++filt',f],capture_output=True,text=True).stdout.strip().replace('(anonymous namespace)::','')[:70]
        print(f'{name:72s}', ' '.join(f'{k}={v}' for k,v in sorted(c.items())))
"
