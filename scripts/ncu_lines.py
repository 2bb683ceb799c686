"""Per-source-line and per-opcode dynamic instruction counts / stall samples from an ncu report's source page.
   python scripts/ncu_lines.py report.ncu-rep [kernel-substring] [top]"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = cur_fn = None; hdr = None
done = set()
i = 0
per = {}
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path": cur_file = r[1]
    elif r and r[0] == "Function Name": cur_fn = r[1]
    elif r and r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr):
        d = per.setdefault(cur_fn, {"lines": collections.OrderedDict(), "ops": collections.Counter(), "opstall": collections.Counter()})
        ix = {h: k for k, h in enumerate(hdr)}
        ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
        if r[0] != "":   # source line row
            key = (cur_file.split("/")[-1], r[0], r[1].strip())
            d["cur"] = key
            d["lines"].setdefault(key, [0, 0])
        else:
            sass = r[3].strip()
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", sass)
            op = m.group(2) if m else "?"
            n = int(r[ie]) if r[ie].isdigit() else 0; s = int(r[isamp]) if r[isamp].isdigit() else 0
            d["ops"][op] += n; d["opstall"][op] += s
            if "cur" in d:
                d["lines"][d["cur"]][0] += n; d["lines"][d["cur"]][1] += s
    i += 1
for fn, d in per.items():
    if want not in fn: continue
    tot = sum(d["ops"].values()); st = sum(d["opstall"].values())
    print(f"=== {fn}: {tot} warp instructions, {st} samples")
    print("  ops:", ", ".join(f"{k} {v*100/tot:.1f}%/{d['opstall'][k]*100/max(st,1):.0f}%s" for k, v in d["ops"].most_common(24)))
    ranked = sorted(d["lines"].items(), key=lambda kv: -kv[1][0])[:top]
    for (f, ln, src), (n, s) in ranked:
        print(f"  {n*100/tot:5.1f}% inst {s*100/max(st,1):5.1f}% stall  {f}:{ln}  {src[:110]}")
