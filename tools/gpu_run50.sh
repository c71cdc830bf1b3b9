O=gpurun_out/r03z; mkdir -p $O
HB_DEBUG_TIMELINE=1 python bench.py --steps 8 --warmup 4 --no-cpu-baseline > $O/bench_tl.json 2> $O/bench_tl.err
grep "hb timeline" $O/bench_tl.err | head -48 > $O/timeline.txt
python - <<'PY'
import re
ev=[]
for l in open('gpurun_out/r03z/timeline.txt'):
    m=re.match(r"\[hb timeline\] rank 0 (land|river) (begin|end) at ([\d.]+) ms", l)
    if m: ev.append((m.group(1), m.group(2), float(m.group(3))))
land=[(b[2],e[2]) for b,e in zip([x for x in ev if x[0]=='land' and x[1]=='begin'],[x for x in ev if x[0]=='land' and x[1]=='end'])]
river=[(b[2],e[2]) for b,e in zip([x for x in ev if x[0]=='river' and x[1]=='begin'],[x for x in ev if x[0]=='river' and x[1]=='end'])]
for i,(l,r) in enumerate(zip(land,river)):
    print(i, 'land %.1f-%.1f (%.1f)'%(l[0],l[1],l[1]-l[0]), 'river %.1f-%.1f (%.1f)'%(r[0],r[1],r[1]-r[0]), 'river starts %.1f after land end'%(r[0]-l[1]))
PY
