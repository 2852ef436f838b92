"""Digest of an `ncu --page source --csv` export: instruction mix by opcode (executed warp instructions, stall
samples) and the top stall sites.  usage: ncu -i rep --page source --csv --kernel-name regex:NAME > x.csv; python tools/ncu_source_digest.py x.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {n: i for i, n in enumerate(hdr)}
ops = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
sites = []
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
stall_tot = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ci["Instructions Executed"]].isdigit(): continue
    src = r[ci["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src
    key = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG")) else op.split(".")[0]
    n = int(r[ci["Instructions Executed"]] or 0); s = int(r[ci["# Samples"]] or 0)
    ops[key][0] += n; ops[key][1] += s; tot_i += n; tot_s += s
    sites.append((s, n, src[:70], {c: int(r[ci[c]] or 0) for c in stall_cols if int(r[ci[c]] or 0)}))
    for c in stall_cols: stall_tot[c] += int(r[ci[c]] or 0)
print("warp instructions %d, samples %d" % (tot_i, tot_s))
for k, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:24]:
    print("%-14s %10d %5.1f%% instr   %5.1f%% samples" % (k, n, 100.0 * n / tot_i, 100.0 * s / max(tot_s, 1)))
print("stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot_s, 1)) for k, v in stall_tot.most_common(8)))
print("top stall sites:")
for s, n, src, st in sorted(sites, key=lambda x: -x[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print("%6d samples %8d exec  %-70s %s" % (s, n, src, " ".join("%s=%d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])))
