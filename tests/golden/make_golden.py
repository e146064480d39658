#!/usr/bin/env python
"""Extract the numeric tables the reference holds for this path into small JSON fixtures.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Sources:
  dune/fem/quadrature/gausspoints_implementation.hh:12-160   1-D Gauss points/weights on [0,1], orders
  dune/fem/space/shapefunctionset/legendrepolynomials.cc:11-27   monomial factors + weights
The reference has no input/output golden vectors for the operator apply; these tables are
the only hard numbers it pins for the path, so the oracle (which regenerates them) is checked against them.
"""
import json, os, re
REF = "/root/reference/dune/fem"
out = os.path.dirname(os.path.abspath(__file__))

src = open(f"{REF}/quadrature/gausspoints_implementation.hh").read()
gauss = {}
m = None
for line in src.splitlines():
    mm = re.match(r"\s*m = (\d+);", line)
    if mm:
        m = int(mm.group(1)); gauss[m] = {"x": {}, "w": {}, "order": None}; continue
    g = re.match(r"\s*G\[m\]\[(\d+)\] = ([0-9.eE+-]+);", line)
    if g and m: gauss[m]["x"][int(g.group(1))] = float(g.group(2))
    w = re.match(r"\s*W\[m\]\[(\d+)\] = ([0-9.eE+-]+);", line)
    if w and m: gauss[m]["w"][int(w.group(1))] = float(w.group(2))
    o = re.match(r"\s*O\[m\] = (-?\d+);", line)
    if o and m: gauss[m]["order"] = int(o.group(1))
table = {}
for m, d in gauss.items():
    if m == 0 or not d["x"]: continue
    table[str(m)] = {"x": [d["x"][i] for i in range(m)], "w": [d["w"][i] for i in range(m)], "order": d["order"]}
json.dump(table, open(f"{out}/gauss_points.json", "w"), indent=1)

src = open(f"{REF}/space/shapefunctionset/legendrepolynomials.cc").read()
wsec = src[src.index("weight["):src.index("factor[")]
weights = [float(v) for v in re.findall(r"[-+]?\d+\.\d+", wsec[wsec.index("="):])]
fsec = src[src.index("factor["):]
fsec = fsec[fsec.index("="):]
rows = re.findall(r"\{([^{}]*)\}", fsec)
factors = [[float(v.replace(" ", "")) for v in r.split(",")] for r in rows]
json.dump({"weight": weights, "factor": factors}, open(f"{out}/legendre_table.json", "w"), indent=1)
print("gauss rules:", sorted(int(k) for k in table), "legendre rows:", len(factors), "weights:", len(weights))
