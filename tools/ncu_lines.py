#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals from an ncu report with -lineinfo.
usage: python tools/ncu_lines.py report.ncu-rep [top N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
cur_file = None
rows = []
hdr = None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] in ("Function Name", "Kernel Name") or hdr is None:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    if len(r) < 10:
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)  # "Source" appears twice: keep the CUDA one
    try:
        inst = int(d["Instructions Executed"])
        samp = int(d["# Samples"])
        thr = int(d["Thread Instructions Executed"])
    except (KeyError, ValueError):
        continue
    rows.append((inst, samp, thr, cur_file, line, d["Source"].strip()[:90]))
ti = sum(r[0] for r in rows) or 1
ts = sum(r[1] for r in rows) or 1
print(f"total warp-instr {ti}, samples {ts}")
for inst, samp, thr, f, line, src in sorted(rows, reverse=True)[:top]:
    print(f"{100*inst/ti:5.1f}% inst {100*samp/ts:5.1f}% samp  act={thr/max(inst,1):4.1f}  {f}:{line}  {src}")
