"""Dev tool: per-source-line summary of an ncu report's source page (samples, instructions, stalls, L1 traffic)."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; f = None; L = []; stall = collections.Counter(); per = {}
for r in rows:
    if r and r[0] == 'File Path': f = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and r and r[0].isdigit():
        def g(name):
            try: return int(r[hdr.index(name)])
            except Exception: return 0
        d = dict(s=g('# Samples'), i=g('Instructions Executed'), gl=g('L1 Tag Requests Global'), sh=g('L1 Wavefronts Shared'), txt=r[1][:75])
        L.append((f, int(r[0]), d))
        c = collections.Counter()
        for j, h in enumerate(hdr):
            if h.startswith('stall_') and 'Not Issued' not in h:
                try: c[h[6:]] += int(r[j])
                except Exception: pass
        stall.update(c); per[(f, int(r[0]))] = c
ts = sum(x[2]['s'] for x in L); ti = sum(x[2]['i'] for x in L); tg = sum(x[2]['gl'] for x in L); tw = sum(x[2]['sh'] for x in L)
print(f"samples {ts}  warp-inst {ti}  L1 global tag req {tg}  shared wavefronts {tw}")
ss = sum(stall.values()); print("stalls: " + "  ".join(f"{k} {v/ss*100:.1f}%" for k, v in stall.most_common(9)))
for x in sorted(L, key=lambda x: -x[2]['s'])[:topn]:
    d = x[2]; c = per[(x[0], x[1])]; t = max(sum(c.values()), 1)
    print(f"{x[0][:14]:14s}{x[1]:4d} s {d['s']/ts*100:5.1f}% i {d['i']/ti*100:5.1f}% gl {d['gl']/max(tg,1)*100:5.1f}% sh {d['sh']/max(tw,1)*100:5.1f}% [{' '.join(f'{k}={v/t*100:.0f}' for k, v in c.most_common(3))}] {d['txt']}")
