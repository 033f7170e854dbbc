#!/usr/bin/env python
"""Static SASS opcode histogram per kernel of libcgbk.so (cuobjdump -sass): the tracked evidence of
what the kernels are made of (LDS widths, MUFU, ATOMS/RED, no tensor-core or TMA opcodes), and a
quick way to compare instruction counts of kernel variants before spending GPU time.

    python tools/sass_hist.py [--lib path] [--match regex] [--top 14] [--json out.json]
"""
import argparse
import collections
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "cgb_b200", "_lib", "libcgbk.so"))
    ap.add_argument("--match", default=".")
    ap.add_argument("--top", type=int, default=14)
    ap.add_argument("--json")
    a = ap.parse_args()
    out = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op = m.group(1)
            parts = op.split(".")
            key = parts[0]
            if parts[0] in ("LDS", "STS", "LDG", "STG", "MUFU", "ATOMS", "RED", "ATOMG", "LDL", "STL", "IMAD", "SHF"):
                key = ".".join(parts[:2]) if len(parts) > 1 else parts[0]
                if parts[0] in ("LDG", "STG", "LDS", "STS", "LDL", "STL"):
                    w = [x for x in parts if x in ("64", "128", "U8", "U16")]
                    key = parts[0] + ("." + w[0] if w else "")
            cur[key] += 1
    rx = re.compile(a.match)
    res = {}
    for name, c in kernels.items():
        if not rx.search(name):
            continue
        tot = sum(c.values())
        res[name] = {"instructions": tot, "opcodes": dict(c.most_common())}
        short = re.sub(r"\(.*", "", name)
        print("%-60s %7d  %s" % (short[:60], tot, " ".join("%s:%d" % kv for kv in c.most_common(a.top))))
    tensor = [k for k, v in res.items() if any(o.startswith(("UTC", "UTMA", "LDTM", "HMMA", "IMMA", "UTMALDG")) for o in v["opcodes"])]
    print("kernels with tensor-core / TMA opcodes:", tensor or "none")
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
