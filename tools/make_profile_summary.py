#!/usr/bin/env python
"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under
profiles/:

    python tools/make_profile_summary.py --tag r1 --rep gpurun_out/prof_conv_r1.ncu-rep \
        --launches gpurun_out/launches_r1.csv

  profiles/<tag>_conv3_1025_b64_ncu.md      key metrics of the `ncu --set full` capture of k_conv_stream
  profiles/<tag>_conv3_1025_b64_raw.csv     the raw page of that capture (one kernel)
  profiles/<tag>_launches.csv               the launch list of `bench.py --steps 1 --warmup 3` (ncu, time only)
  profiles/<tag>_launch_shares.md           per-kernel share of the profiled launches
  profiles/traffic_conv3_1025_b64.json      DRAM bytes per launch (read by bench.py -> roofline.traffic)
"""
import argparse
import collections
import csv
import io
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    ("gpu__time_duration.sum", "kernel time (cold, serialised under ncu)"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("launch__grid_size", "CTAs (one warp each)"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / scheduler"),
    ("sm__cycles_active.avg", "SM cycles active (avg)"),
]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * scale.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--bench")
    args = ap.parse_args()
    os.makedirs(PROF, exist_ok=True)
    if args.rep:
        raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, r = rows[0], rows[1], rows[2]
        with open(os.path.join(PROF, "%s_conv3_1025_b64_raw.csv" % args.tag), "w") as fh:
            fh.write(raw)
        get = lambda k: (r[hdr.index(k)], units[hdr.index(k)]) if k in hdr else ("n/a", "")
        lines = ["# ncu --set full: %s" % get("Kernel Name")[0].strip(), "",
                 "Capture: `ncu --set full -k regex:k_conv_wide -s 150 -c 1 --clock-control none --import-source on "
                 "python bench.py --steps 1 --warmup 3 --no-cpu --no-ttr` on one B200 (workload conv3_jacobi_square_1025_b64: "
                 "64 x 1025^2, clamp, one Phi_H application per launch).", "", "| metric | value |", "|---|---|"]
        for k, label in KEYS:
            v, u = get(k)
            lines.append("| %s (`%s`) | %s %s |" % (label, k, v, u))
        lines += ["", "Warp stall reasons (warps per issue-active cycle):", "", "| reason | value |", "|---|---|"]
        stalls = []
        for i, k in enumerate(hdr):
            if "issue_stalled" in k and "per_issue_active" in k:
                try:
                    stalls.append((float(r[i].replace(",", "")), k.split("issue_stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
        for v, name in sorted(stalls, reverse=True):
            if v >= 0.01:
                lines.append("| %s | %.3f |" % (name, v))
        rd = to_bytes(*get("dram__bytes_read.sum"))
        wr = to_bytes(*get("dram__bytes_write.sum"))
        algo = 8.0 * 64 * 1025 * 1025
        lines += ["", "DRAM traffic per launch: %.1f MB read + %.1f MB written = %.1f MB; algorithmic bytes "
                  "(8 B per cell-update x 64 x 1025^2) = %.1f MB, ratio %.3f." % (rd / 1e6, wr / 1e6, (rd + wr) / 1e6,
                                                                                  algo / 1e6, (rd + wr) / algo)]
        with open(os.path.join(PROF, "%s_conv3_1025_b64_ncu.md" % args.tag), "w") as fh:
            fh.write("\n".join(lines) + "\n")
        with open(os.path.join(PROF, "traffic_conv3_1025_b64.json"), "w") as fh:
            json.dump({"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "algorithmic_bytes_per_launch": algo,
                       "source": "profiles/%s_conv3_1025_b64_ncu.md (ncu --set full, one launch)" % args.tag}, fh, indent=1)
    if args.launches:
        dst = os.path.join(PROF, "%s_launches.csv" % args.tag)
        with open(args.launches) as fh:
            text = [l for l in fh if not l.startswith("==")]
        with open(dst, "w") as fh:
            fh.writelines(text)
        rows = list(csv.DictReader(io.StringIO("".join(text))))
        tot = collections.defaultdict(lambda: [0, 0.0])
        for row in rows:
            if row.get("Metric Name") != "gpu__time_duration.sum":
                continue
            name = row["Kernel Name"].split("(")[0].replace("void ", "").strip()
            t = float(row["Metric Value"].replace(",", ""))
            if row["Metric Unit"].startswith("ns"):
                t /= 1e3
            elif row["Metric Unit"].startswith("ms"):
                t *= 1e3
            tot[name][0] += 1
            tot[name][1] += t
        total = sum(v[1] for v in tot.values())
        lines = ["# launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -c 320 python bench.py "
                 "--steps 1 --warmup 3 --no-cpu --no-ttr`", "",
                 "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
                 "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
        for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            lines.append("| %s | %d | %.1f | %.2f | %.1f%% |" % (name[-70:], n, t, t / n, 100 * t / total))
        with open(os.path.join(PROF, "%s_launch_shares.md" % args.tag), "w") as fh:
            fh.write("\n".join(lines) + "\n")
    if args.bench:
        shutil.copy(args.bench, os.path.join(PROF, "%s_bench.json" % args.tag))


if __name__ == "__main__":
    main()
