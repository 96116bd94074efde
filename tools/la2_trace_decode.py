"""Decode the LA2TRACE lines of LDM_LA2_TRACE=1 (linattn_tc2_kernel<true>): per-event clock deltas for one sample."""
import re, sys
ev = {}
for l in open(sys.argv[1]):
    m = re.match(r'LA2TRACE w(\d+) (\d+) (\d+)', l)
    if m: ev.setdefault(int(m.group(1)), []).append((int(m.group(2)), int(m.group(3))))
def unwrap(seq):
    out=[]; base=0; prev=None
    for tag,c in seq:
        if prev is not None and c < prev: base += 1<<24
        prev=c; out.append((tag,c+base))
    return out
for w in ev: ev[w]=unwrap(ev[w])
epi = sorted(w for w in ev if w != 1)
samples = [i for i,(t,c) in enumerate(ev[epi[0]]) if t==10]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t_start = ev[epi[0]][samples[which]][1]
t_end = ev[epi[0]][samples[which+1]][1] if which+1 < len(samples) else ev[epi[0]][-1][1]
print("sample cycles", t_end - t_start)
for w in sorted(ev):
    seq=[(t,c) for t,c in ev[w] if t_start-500 <= c <= t_end]
    prev=seq[0][1]; line=[]
    for t,c in seq:
        line.append(f"{t}:{c-prev}"); prev=c
    print("warp", w, " ".join(line))
