#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: executed-instruction histogram by opcode weighted by
execution count, hot instruction footprint (bytes of code executed >= 1% as often as the hottest
instruction) and the top stall reasons. usage: ncu_src.py file.csv [passes]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
ex = [int(r[col["Instructions Executed"]] or 0) for r in data]
mx = max(ex)
hot = [r for r, e in zip(data, ex) if e >= 0.5 * mx]
warm = [r for r, e in zip(data, ex) if e >= 0.01 * mx]
print("instructions in kernel %d, executed at >=50%% of max: %d (%.1f KB), >=1%% of max: %d (%.1f KB)" % (
    len(data), len(hot), len(hot) * 16 / 1024, len(warm), len(warm) * 16 / 1024))
tot = sum(ex)
c = collections.Counter()
for r, e in zip(data, ex):
    t = re.sub(r'^@!?U?P\d+\s+', '', r[col["Source"]].strip())
    op = t.split()[0] if t else "?"
    op = "MOV*" if op.startswith("IMAD.MOV") or op == "MOV" else op.split(".")[0]
    c[op] += e
print("warp-instructions executed %.3e; per max-count line: %.0f" % (tot, tot / mx))
print(" ".join("%s:%.1f%%" % (k, 100 * v / tot) for k, v in c.most_common(25)))
st = collections.Counter()
for h in hdr:
    if h.startswith("stall_") and "Not Issued" not in h:
        st[h] = sum(int(r[col[h]] or 0) for r in data)
ts = sum(st.values())
print(" ".join("%s:%.1f%%" % (k, 100 * v / ts) for k, v in st.most_common(10)))
