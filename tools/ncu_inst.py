"""Dev tool: per-source-line instruction shares of an ncu report (source page), sorted by executed instructions."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; f = None; L = []
for r in rows:
    if r and r[0] == 'File Path': f = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and r and r[0].isdigit():
        def g(name):
            try: return int(r[hdr.index(name)])
            except Exception: return 0
        L.append((f, int(r[0]), g('Instructions Executed'), g('# Samples'), r[1][:110]))
ti = sum(x[2] for x in L); ts = sum(x[3] for x in L)
acc = 0
for x in sorted(L, key=lambda x: -x[2])[:topn]:
    acc += x[2]
    print(f"{x[0][:14]:14s}{x[1]:4d} i {x[2]/ti*100:5.1f}% (cum {acc/ti*100:5.1f}%) s {x[3]/ts*100:5.1f}%  {x[4]}")
