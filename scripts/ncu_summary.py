"""Markdown summary of an `ncu --set full` report: per captured kernel the headline metrics (raw page) and the warp
stall breakdown + hottest SASS lines (source page).  usage: python scripts/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
ix = {h: i for i, h in enumerate(hdr)}
print(f"# ncu --set full summary: `{rep.split('/')[-1]}`\n")
print("| metric | " + " | ".join(f"`{r[ix['Kernel Name']].split('(')[0][-28:]}`" for r in data) + " |")
print("|---|" + "---:|" * len(data))
for m in want:
    if m in ix:
        print(f"| {m} ({units[ix[m]]}) | " + " | ".join(r[ix[m]] for r in data) + " |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur); continue
    if cur is not None:
        cur["rows"].append(r)
for b in blocks:
    h, d = b["rows"][0], b["rows"][1:]
    jx = {x: i for i, x in enumerate(h)}
    stall = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    tot = {x: 0 for x in stall}; S = 0
    for r in d:
        try:
            S += int(r[jx["# Samples"]])
            for x in stall: tot[x] += int(r[jx[x]])
        except Exception:
            pass
    print(f"\n## `{b['name'].split('(')[0]}`: warp stall samples ({S})\n")
    print(", ".join(f"{x[6:]} {100 * v / max(S, 1):.1f}%" for x, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
    top = sorted(d, key=lambda r: -int(r[jx["# Samples"]]) if r[jx["# Samples"]].isdigit() else 0)[:8]
    print("\nhottest SASS lines (samples, instruction, stall reasons):\n")
    for r in top:
        print(f"- {r[jx['# Samples']]} `{r[jx['Source']].strip()[:80]}` " + ", ".join(f"{x[6:]}={r[jx[x]]}" for x in stall if r[jx[x]] not in ("0", "")))
