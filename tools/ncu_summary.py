#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, without a GPU) as the markdown table kept under profiles/ and
as the per-kernel DRAM traffic file bench.py cross-references (profiles/traffic.json).

    python tools/ncu_summary.py gpurun_out/final_prof.ncu-rep --title "..." --md profiles/ncu_full_r02_final.md \
        --traffic profiles/traffic.json
"""
import argparse
import collections
import csv
import io
import json
import re
import subprocess

STALLS = ("long_scoreboard", "short_scoreboard", "barrier", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "no_instruction", "not_selected", "branch_resolving", "dispatch_stall", "membar", "drain", "imc_miss", "sleeping",
          "tex_throttle", "misc")


def short_name(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("wefax::", "").replace("fast::", "")
    return name


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--title", default="ncu --set full")
    ap.add_argument("--md")
    ap.add_argument("--traffic")
    ap.add_argument("--command", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, key, default=0.0):
        i = col.get(key)
        if i is None or r[i] in ("", "n/a"):
            return default
        try:
            return float(r[i].replace(",", ""))
        except ValueError:
            return default

    out = []
    traffic = collections.defaultdict(list)
    for r in rows[2:]:
        name = short_name(r[col["Kernel Name"]])
        stalls = sorted(((get(r, f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"), s) for s in STALLS), reverse=True)[:3]
        rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
        # units differ per column (Mbyte / Kbyte / byte): normalise through the unit row
        def mb(key, v):
            unit = rows[1][col[key]].lower() if key in col else "byte"
            return v * {"gbyte": 1e3, "mbyte": 1.0, "kbyte": 1e-3, "byte": 1e-6}.get(unit, 1e-6)
        rd, wr = mb("dram__bytes_read.sum", rd), mb("dram__bytes_write.sum", wr)
        dur = get(r, "gpu__time_duration.sum")
        unit = rows[1][col["gpu__time_duration.sum"]].lower()
        dur_us = dur * {"us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(unit, 1.0)
        traffic[re.sub(r"<.*$", "", name)].append((rd + wr) * 1e6)
        out.append(dict(name=name, us=dur_us, rd=rd, wr=wr, dram=get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        issue=get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        warps=get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                        regs=int(get(r, "launch__registers_per_thread")), inst=get(r, "smsp__inst_executed.sum") / 1e6,
                        stalls=", ".join(f"{s} {v:.1f}" for v, s in stalls)))
    lines = [f"# {args.title}", ""]
    if args.command:
        lines += [f"Command: `{args.command}`", ""]
    lines += ["(per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes)", "",
              "| kernel | time us | DRAM read MB | DRAM write MB | DRAM % | issue % | warps active % | regs | warp insts (M) | top stalls (warp-cycles per issue) |",
              "|---|---|---|---|---|---|---|---|---|---|"]
    for o in out:
        lines.append(f"| {o['name']} | {o['us']:.1f} | {o['rd']:.1f} | {o['wr']:.1f} | {o['dram']:.1f} | {o['issue']:.1f} | "
                     f"{o['warps']:.1f} | {o['regs']} | {o['inst']:.1f} | {o['stalls']} |")
    total_us = sum(o["us"] for o in out)
    total_mb = sum(o["rd"] + o["wr"] for o in out)
    lines += ["", f"Sum over the {len(out)} launches: {total_us:.0f} us, {total_mb / 1e3:.2f} GB of DRAM traffic."]
    text = "\n".join(lines) + "\n"
    if args.md:
        with open(args.md, "w") as fh:
            fh.write(text)
    else:
        print(text)
    if args.traffic:
        with open(args.traffic, "w") as fh:
            blob = {k: sum(v) / len(v) for k, v in traffic.items()}
            blob["unit"] = ("bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, averaged over "
                            "the launches of one step)")
            blob["source"] = args.md or args.report
            json.dump(blob, fh, indent=1)


if __name__ == "__main__":
    main()
