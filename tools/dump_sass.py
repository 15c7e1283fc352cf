#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel opcode histogram plus the full listing of the kernels
named on the command line.  Usage: python tools/dump_sass.py <tag> (writes profiles/sass_<tag>_*.txt)"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "sage-3d_official_b200" / "libsage_splat.so"
GROUPS = {"composite": ["k_composite"], "sort": ["k_sort_hist", "k_sort_scan", "k_sort_scatter", "k_sort_fused"],
          "emit": ["k_emit", "k_bin_order"], "project": ["k_project", "k_sh_color"]}


def main(tag):
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = cur.replace("(anonymous namespace)::", "")
            funcs[cur] = []
        elif cur is not None and re.search(r"/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(line.rstrip())
    for group, names in GROUPS.items():
        out = [f"# cuobjdump -sass {LIB.name} (sm_100a), kernels of group '{group}'", ""]
        for fn, lines in funcs.items():
            if not any(fn.startswith(n) or f" {n}" in fn.split("(")[0] for n in names):
                continue
            ops = collections.Counter()
            for l in lines:
                m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
                if m:
                    ops[m.group(1)] += 1
            out.append(f"## {fn.split('(')[0]}: {sum(ops.values())} instructions")
            keys = ("FFMA2", "FMUL2", "FADD2", "LDGSTS", "UBLKCP", "SYNCS", "LDGDEPBAR", "DEPBAR", "MATCH", "VOTE", "ATOMS", "ATOMG", "RED", "LDS", "STS", "LDG", "STG", "BAR", "MUFU", "SHFL", "POPC")
            fam = collections.Counter()
            for op, c in ops.items():
                for k in keys:
                    if op.startswith(k):
                        fam[k] += c
                        break
            out.append("   families: " + ", ".join(f"{k} x{fam[k]}" for k in keys if fam[k]))
            out.append("   top opcodes: " + ", ".join(f"{o} x{c}" for o, c in ops.most_common(24)))
            out.append("")
            if group in ("composite",) or "k_sort_scatter<8, false, 16, true>" in fn or fn.startswith("k_emit") or "k_sort_fused" in fn:
                out += lines + [""]
        (ROOT / "profiles" / f"sass_{tag}_{group}.txt").write_text("\n".join(out) + "\n")
        print(f"profiles/sass_{tag}_{group}.txt: {len(out)} lines")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
