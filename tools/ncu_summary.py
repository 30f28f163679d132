"""Summarise an .ncu-rep into markdown: python tools/ncu_summary.py REPORT.ncu-rep "title" > profiles/x.md"""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# {title}\n")
print(f"Source: `{rep}` (ncu --set full --clock-control none --import-source on).  Numbers under a profiler are for SHARES "
      "and counters only; timings quoted elsewhere come from CUDA events in un-profiled runs.\n")
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"## `{name}`\n\n| metric | value | unit |\n|---|---|---|")
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            print(f"| {h} | {v} | {u} |")
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
    # several kernels in one report repeat the header block: keep the well-formed data rows only
    data = [r for r in data if len(r) == len(hdr) and (r[ix["# Samples"]] or "0").isdigit()]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
    print("## Warp-stall samples by reason\n\n| reason | samples | share |\n|---|---|---|")
    agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall}
    for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        print(f"| {h} | {v} | {100 * v / tot:.1f}% |")
    print("\n## Hottest SASS instructions (by samples)\n\n| samples | executed | instruction | top stall |\n|---|---|---|---|")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
        st = max(((int(r[ix[h]] or 0), h) for h in stall))
        print(f"| {r[ix['# Samples']]} | {r[ix['Instructions Executed']]} | `{r[ix['Source']].strip()[:60]}` | {st[1]} ({st[0]}) |")
