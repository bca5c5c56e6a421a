"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the last train step (from the last
enc_fwd_fused / embed_ln_fwd launch to the end).  usage: python scripts/launch_summary.py file.csv [--md]"""
import csv, re, sys
rows = []
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 10 and r[0].isdigit():
        rows.append((int(r[0]), r[4], r[7], r[8], float(r[-1])))
names = [re.sub(r'\(.*', '', x[1]).replace('void ', '').replace('b4r::', '') for x in rows]
idx = [i for i, n in enumerate(names) if 'enc_fwd_fused' in n or 'embed_ln_fwd' in n]
s = idx[-1]
if s > 0 and 'mlm_select' in names[s - 1]:   # the slot selection is enqueued first (parallel branch beside the forward)
    s -= 1
tot = sum(r[4] for r in rows[s:])
md = '--md' in sys.argv
if md:
    print("| # | kernel | block | grid | us | share |\n|---:|---|---|---|---:|---:|")
for i in range(s, len(rows)):
    if md:
        print(f"| {i - s} | `{names[i][:58]}` | {rows[i][2]} | {rows[i][3]} | {rows[i][4] / 1e3:.1f} | {100 * rows[i][4] / tot:.1f}% |")
    else:
        print(f"{i - s:3d} {names[i][:58]:58s} {rows[i][2]:>14s} {rows[i][3]:>14s} {rows[i][4] / 1e3:8.1f} us {100 * rows[i][4] / tot:5.1f}%")
print(("\n" if md else "") + f"total {tot / 1e3:.1f} us over {len(rows) - s} launches")
