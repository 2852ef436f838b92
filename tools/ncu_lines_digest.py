"""Per-CUDA-line digest of `ncu -i rep --page source --csv --print-source cuda,sass --kernel-id :::N`: executed warp
instructions and stall samples per source line (inlined code is attributed to the line it came from)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    ie = r[hdr["Instructions Executed"]]
    if not ie.isdigit(): continue
    sm = r[hdr["# Samples"]]
    out.append((int(ie), int(sm) if sm.isdigit() else 0, cur_file, int(r[0]), r[1].strip()[:110]))
ti = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print("total warp instructions %d, samples %d" % (ti, ts))
for ie, s, f, ln, src in sorted(out, key=lambda o: -o[0])[:top]:
    print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100.0 * ie / ti, 100.0 * s / max(ts, 1), f, ln, src))
