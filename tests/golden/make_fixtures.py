#!/usr/bin/env python3
"""Generates tests/golden/*.json from the reference's own example data and test files.

Run in the build container (where /root/reference exists):  python tests/golden/make_fixtures.py
The GPU box has no /root/reference, so the derived fixtures are committed.

What is taken from the reference (data and known answers only, no source code):
  * exampleData/fourPop.histogram.txt            -> patterns / counts / N / TOTAL_SITES
  * exampleData/fourPop.template.txt             -> instantiated event list (ModelTemplate.hs:176-213 rules)
  * exampleData/fourPop.m4.mcmc.paramEstimates.txt:4-21 -> MaxL column + Score (whole-path golden, -logl)
  * src/RarecoalLib/Core/Test.hs:129-135,146,177-181    -> 5-pop fixture + getProb known answers
  * src/RarecoalLib/StateSpace/Test.hs:82-95            -> fillUpStateSpace known answer
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

EV_JOIN, EV_SPLIT, EV_POPSIZE, EV_FREEZE = 0, 1, 2, 3


def parse_template(text):
    br = re.search(r"BRANCHES\s*=\s*\[(.*?)\]", text, re.S).group(1)
    branches = [b.strip() for b in br.split(",")]
    ev = re.search(r"EVENTS\s*=\s*\[(.*?)\]", text, re.S).group(1)
    cmds = [c.split() for c in ev.split(";") if c.strip()]
    return branches, cmds


def instantiate(branches, cmds, params):
    """ModelTemplate.hs:176-213: K emits Join then SetPopSize at the same time."""
    def val(tok):
        return params[tok[1:-1]] if tok.startswith("<") else float(tok)
    idx = {b: i for i, b in enumerate(branches)}
    events = []
    for c in cmds:
        kind = c[0]
        if kind == "P":
            events.append([val(c[1]), EV_POPSIZE, idx[c[2]], 0, val(c[3])])
        elif kind == "J":
            events.append([val(c[1]), EV_JOIN, idx[c[2]], idx[c[3]], 0.0])
        elif kind == "K":
            events.append([val(c[1]), EV_JOIN, idx[c[2]], idx[c[3]], 0.0])
            events.append([val(c[1]), EV_POPSIZE, idx[c[2]], 0, val(c[4])])
        elif kind == "S":
            events.append([val(c[1]), EV_SPLIT, idx[c[2]], idx[c[3]], val(c[4])])
        elif kind == "F":
            events.append([val(c[1]), EV_FREEZE, idx[c[2]], 0, 1.0 if c[3] == "True" else 0.0])
        else:
            raise ValueError(c)
    return events


def main():
    ex = os.path.join(REF, "exampleData")
    lines = open(os.path.join(ex, "fourPop.histogram.txt")).read().split("\n")
    names = lines[0].split("=")[1].split(",")
    nvec = [int(x) for x in lines[1].split("=")[1].split(",")]
    max_m = int(lines[2].split("=")[1])
    total = int(lines[3].split("=")[1])
    rows = [l.split() for l in lines[4:] if l.strip()]
    pats = [[int(x) for x in r[0].split(",")] for r in rows]
    cnts = [int(r[1]) for r in rows]

    plines = open(os.path.join(ex, "fourPop.m4.mcmc.paramEstimates.txt")).read().split("\n")
    params, score = {}, None
    for l in plines[3:]:
        w = l.split()
        if not w:
            continue
        if w[0] == "Score":
            score = float(w[1])
        else:
            params[w[0]] = float(w[1])
    branches, cmds = parse_template(open(os.path.join(ex, "fourPop.template.txt")).read())
    events = instantiate(branches, cmds, params)
    json.dump({
        "source": "exampleData/fourPop.{histogram,template}.txt + fourPop.m4.mcmc.paramEstimates.txt (MaxL column)",
        "names": names, "branches": branches, "nVec": nvec, "maxM": max_m, "totalSites": total,
        "patterns": pats, "counts": cnts, "params": params, "events": events, "theta": 0.0005,
        "score_m4_maxl": score,
        "note": "score = -logl at maxAf=4, theta=0.0005, regularization 0 (SURVEY.md F5)",
    }, open(os.path.join(OUT, "fourpop.json"), "w"))

    json.dump({
        "source": "src/RarecoalLib/Core/Test.hs:129-135,146,177-181",
        "nrPops": 5, "theta": 0.0005, "nVec": [100] * 5, "disc": [1.0] * 5,
        "events": [[0.0025, EV_JOIN, 0, 1, 0.0], [0.006, EV_JOIN, 2, 3, 0.0],
                   [0.0075, EV_JOIN, 2, 4, 0.0], [0.01, EV_JOIN, 0, 2, 0.0]],
        "kat": [
            {"config": [0, 1, 2, 0, 1], "prob": 2.009581497800885e-6, "live_in_ref_test": True},
            {"config": [1, 1, 1, 1, 1], "prob": 1.5455236177098954e-6, "live_in_ref_test": False},
            {"config": [2, 3, 0, 0, 0], "prob": 7.503194796424192e-6, "live_in_ref_test": False},
            {"config": [1, 0, 2, 3, 2], "prob": 4.355291335648456e-7, "live_in_ref_test": False},
            {"config": [0, 1, 0, 0, 1], "prob": 1.2611453629387214e-5, "live_in_ref_test": True},
        ],
        "fillup": {"source": "src/RarecoalLib/StateSpace/Test.hs:82-95", "nrPops": 5, "maxAf": 4,
                   "seeds": [[2, 2, 0, 0, 0], [0, 2, 2, 0, 0]],
                   "expected": [[1, 1, 0, 0, 0], [1, 2, 0, 0, 0], [2, 1, 0, 0, 0], [2, 2, 0, 0, 0],
                                [0, 1, 1, 0, 0], [0, 1, 2, 0, 0], [0, 2, 1, 0, 0], [0, 2, 2, 0, 0]]},
    }, open(os.path.join(OUT, "fivepop_kat.json"), "w"))
    # the example DATA files themselves (formats of SURVEY.md Appendix B), for the front-end tests
    import shutil
    for name in ("fourPop.template.txt", "fourPop.histogram.txt", "fourPop.m4.mcmc.paramEstimates.txt"):
        shutil.copyfile(os.path.join(ex, name), os.path.join(OUT, name))
    print("wrote fourpop.json, fivepop_kat.json and the three example data files")


if __name__ == "__main__":
    main()
