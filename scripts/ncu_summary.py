"""Turn an ncu report (.ncu-rep) into the per-kernel summary committed under profiles/.
   python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/name [--traffic profiles/traffic.json]
Writes profiles/name.md (one table row per captured launch + the stall mix of the source page) and, with
--traffic, updates the dram bytes per launch of each kernel (what bench.py reports as roofline.traffic)."""
import csv, io, json, os, subprocess, sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 pipe %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def short(n):
    if "fwd_step_kernel" in n:
        return "fwd_step_kernel<save_frames>" if ("<1>" in n or "(bool)1" in n) else "fwd_step_kernel<no_frames>"
    for k in ("rev_image_kernel", "adj_step_kernel"):
        if k in n:
            return k
    return n.split("(")[0][-40:]


def main():
    rep, outbase = sys.argv[1], sys.argv[2]
    traffic_path = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary of `{os.path.basename(rep)}`", "",
             "`ncu --set full --clock-control none --import-source on`, one row per captured launch "
             "(cold-cache, serialised replays: compare shares, not absolutes).", "",
             "| kernel | " + " | ".join(t for _, t in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
    traffic = json.load(open(traffic_path)) if traffic_path and os.path.exists(traffic_path) else {}
    seen = {}
    for r in data:
        name = short(r[ix["Kernel Name"]])
        cells = []
        for m, _ in METRICS:
            if m in ix:
                v, u = r[ix[m]], units[ix[m]]
                try:
                    f = float(v)
                    cells.append(f"{f:.4g} {u}".strip())
                except ValueError:
                    cells.append(v)
            else:
                cells.append("-")
        lines.append(f"| {name} | " + " | ".join(cells) + " |")
        try:
            rd = float(r[ix["dram__bytes_read.sum"]]) * SCALE.get(units[ix["dram__bytes_read.sum"]], 1.0)
            wr = float(r[ix["dram__bytes_write.sum"]]) * SCALE.get(units[ix["dram__bytes_write.sum"]], 1.0)
            seen.setdefault(name, []).append(rd + wr)
        except Exception:
            pass
    # stall mix per kernel from the source page
    lines += ["", "## warp stall mix (source page, all samples)", ""]
    names = []
    for r in data:
        n = r[ix["Kernel Name"]]
        if n not in names:
            names.append(n)
    for n in names:
        key = n.split("(")[0].split("::")[-1].split("<")[0]
        src = page(rep, "source", ["--kernel-name", f"regex:{key}", "--launch-count", "1"])
        if len(src) < 3:
            continue
        h = src[1]
        sx = {c: i for i, c in enumerate(h)}
        stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[sx["# Samples"]] or 0) for r in src[2:])
        inst = sum(int(r[sx["Instructions Executed"]] or 0) for r in src[2:])
        mix = {c: sum(int(r[sx[c]] or 0) for r in src[2:]) for c in stalls}
        top = sorted(mix.items(), key=lambda kv: -kv[1])[:8]
        lines.append(f"* `{short(n)}`: {inst} warp instructions, {tot} samples: " +
                     ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in top if v))
    open(outbase + ".md", "w").write("\n".join(lines) + "\n")
    if traffic_path:
        for k, v in seen.items():
            traffic[k] = sum(v) / len(v)
        json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
