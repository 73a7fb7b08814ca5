"""Basic-block view of an .ncu-rep: consecutive SASS instructions with the same execution count are merged.
usage: python tools/ncu_blocks.py report.ncu-rep [min_share_percent]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; ins = []
for r in rows:
    if r and r[0] in ("Address", "#"): hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    try: ins.append((d.get("Address"), d.get("Source"), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])))
    except (ValueError, KeyError): pass
tot = sum(i[2] for i in ins) or 1; tots = sum(i[4] for i in ins) or 1
print("instructions %d, executed %d, samples %d" % (len(ins), tot, tots))
blocks = []; cur = None
for n, (a, s, e, t, smp) in enumerate(ins):
    if cur is None or cur["e"] != e: 
        cur = {"start": n, "e": e, "n": 0, "t": 0, "smp": 0, "first": s.strip(), "ops": {}}; blocks.append(cur)
    cur["n"] += 1; cur["t"] += t; cur["smp"] += smp
    op = s.strip().split()[0] if not s.strip().startswith("@") else s.strip().split()[1]
    op = op.split(".")[0]; cur["ops"][op] = cur["ops"].get(op, 0) + 1
for b in blocks:
    share = 100.0 * b["e"] * b["n"] / tot
    if share >= min_pct or 100.0 * b["smp"] / tots >= min_pct:
        top = ",".join("%s%d" % kv for kv in sorted(b["ops"].items(), key=lambda kv: -kv[1])[:6])
        print("@%5d n=%4d exec=%9d inst %5.1f%% samp %5.1f%% lanes %4.1f | %s" % (b["start"], b["n"], b["e"], share, 100.0 * b["smp"] / tots, b["t"] / max(1, b["e"] * b["n"]), top))
