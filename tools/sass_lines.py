#!/usr/bin/env python
"""Instruction count per source line of one kernel of a cubin (nvdisasm -g output on stdin or a file).
    cuobjdump -xelf all <obj>; nvdisasm -g <cubin> > dis.txt; python tools/sass_lines.py dis.txt <substring of the section name> [top]
"""
import collections
import re
import sys


def main():
    path, needle = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    sec, cur, data = None, None, {}
    for l in open(path):
        m = re.match(r'\s*\.section\s+(\S+?),', l)
        if m:
            sec = m.group(1) if m.group(1).startswith(".text.") else None
            if sec:
                data.setdefault(sec, collections.Counter())
            cur = None
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if sec and re.match(r'\s+/\*[0-9a-f]{4,6}\*/', l):
            data[sec][cur] += 1
    for s, c in data.items():
        if needle not in s:
            continue
        tot = sum(c.values())
        print(s[:80], tot, "instructions", round(tot * 16 / 1024, 1), "KB")
        f = collections.Counter()
        for k, v in c.items():
            f[k[0] if k else None] += v
        print(" per file:", f.most_common())
        for k, v in sorted(c.items(), key=lambda x: -x[1])[:top]:
            print("  ", k, v)


if __name__ == "__main__":
    main()
