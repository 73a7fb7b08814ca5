"""Per-source-line totals (instructions executed, stall samples) from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_lines.py report.ncu-rep [min_share_percent]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(raw)))
cur = None; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r or r[0] == "": continue   # SASS rows have an empty line number
    d = dict(zip(hdr, r))
    try: out.append((cur, int(d["Line No"]), int(d["Instructions Executed"]), int(d["# Samples"]), int(d["Thread Instructions Executed"]), r[1][:90]))
    except ValueError: pass
tot_i = sum(o[2] for o in out) or 1; tot_s = sum(o[3] for o in out) or 1
print("total warp-instr %d, samples %d" % (tot_i, tot_s))
for f, ln, ins, smp, tins, src in sorted(out):
    if 100.0 * ins / tot_i >= min_pct or 100.0 * smp / tot_s >= min_pct:
        print("%-12s %4d  inst %5.1f%%  samples %5.1f%%  lanes %4.1f  | %s" % (f, ln, 100.0 * ins / tot_i, 100.0 * smp / tot_s, tins / max(ins, 1), src))
