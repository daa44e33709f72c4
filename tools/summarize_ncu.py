#!/usr/bin/env python
"""Summarises gpurun_out ncu artefacts into profiles/ (run in the build container, no GPU needed).

  python tools/summarize_ncu.py <round-tag> <launches.csv> <full.ncu-rep> [kernel-regex]
"""
import csv
import collections
import io
import re
import subprocess
import sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
out = ["# ncu summary %s" % tag, ""]

# ---- launch list: per-kernel share of the profiled command
rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
out += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare SHARES)", "",
        "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| `%s` | %d | %.3f | %.1f%% |" % (k, n, ns / 1e6, 100 * ns / tot))
out.append("")

# ---- full capture: key counters of the top kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rr[0], rr[1], rr[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpc__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg.per_second"]
out += ["## `ncu --set full` capture of the dominant kernel", "", "| metric | unit | value |", "|---|---|---|"]
for h, u, v in zip(hdr, units, vals):
    base = h.split(".TriageCompute.")[-1]
    if base in want or h in want:
        out.append("| %s | %s | %s |" % (base, u, v))
out.append("")

# ---- stall summary from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sr = list(csv.reader(io.StringIO(src)))
h2 = sr[1]
ix = {h: i for i, h in enumerate(h2)}
data = sr[2:]
tot_s = sum(int(r[ix["# Samples"]] or 0) for r in data)
stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
agg2 = sorted(((sum(int(r[ix[h]] or 0) for r in data), h) for h in stalls), reverse=True)
out += ["## warp-stall samples over the whole kernel (%d samples, %d SASS instructions)" % (tot_s, len(data)), ""]
out += ["| stall | share |", "|---|---:|"] + ["| %s | %.1f%% |" % (h, 100.0 * s / max(tot_s, 1)) for s, h in agg2[:8]]
out.append("")
mn = collections.Counter()
for r in data:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    if m:
        mn[m.group(1).split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
out += ["SASS mnemonics proving the Blackwell path (executed warp-instructions): " +
        ", ".join("%s=%d" % (k, mn[k]) for k in ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS") if k in mn), ""]
out += ["top instructions by samples:", "", "```"]
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
    out.append("%7s  exec=%-10s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]][:90]))
out += ["```", ""]
open("profiles/%s_ncu_summary.md" % tag, "w").write("\n".join(out))
print("\n".join(out))
