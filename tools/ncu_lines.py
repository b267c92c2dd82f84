#!/usr/bin/env python
"""Executed warp instructions per source line: joins the SASS page of an ncu capture (ncu -i X.ncu-rep --page source --csv)
with the line table of the same kernel (nvdisasm -g of the cubin, cuobjdump -xelf all <obj>), instruction by instruction.
    python tools/ncu_lines.py sass_page.csv dis.txt <substring of the .text section name> [top] [units] [source dir]
units: divide the counts by this number (e.g. maps per launch); source dir: print the text of each line from there.
"""
import collections
import csv
import re
import sys


def line_table(path, needle):
    sec, cur, out = None, None, []
    for l in open(path):
        m = re.match(r'\s*\.section\s+(\S+?),', l)
        if m:
            sec = m.group(1) if m.group(1).startswith(".text.") and needle in m.group(1) else None
            cur = None
            continue
        if sec is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', l)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2)))
    return out


def main():
    page, dis, needle = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    units = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
    srcdir = sys.argv[6] if len(sys.argv) > 6 else None
    cache = {}

    def source_text(line):
        if not srcdir or not line:
            return ""
        if line[0] not in cache:
            try:
                cache[line[0]] = open(srcdir + "/" + line[0]).read().splitlines()
            except OSError:
                cache[line[0]] = []
        src = cache[line[0]]
        return "  | " + src[line[1] - 1].strip()[:110] if 0 < line[1] <= len(src) else ""
    rows = list(csv.reader(open(page)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]  # one table per captured launch: take the last
    rows = rows[starts[-1]:]
    hdr = rows[1]
    ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    ins = [(r[isrc], int(r[iex] or 0)) for r in rows[2:] if len(r) > iex]
    table = line_table(dis, needle)
    if len(ins) != len(table):
        print("instruction counts differ: page %d, disassembly %d" % (len(ins), len(table)))
    per_line, per_file = collections.Counter(), collections.Counter()
    total = 0
    mismatch = 0
    for (src, ex), (_, line, text) in zip(ins, table):
        if src.split()[0].strip("@!P0123456789 ") and text.split()[0] != src.split()[0] and src.split()[0][0] != "@":
            mismatch += 1
        per_line[line] += ex
        per_file[line[0] if line else None] += ex
        total += ex
    print("executed warp instructions:", total, " opcode mismatches:", mismatch, (" per unit: %.1f" % (total / units)) if units else "")
    print("per file:", [(k, round(v / units, 1) if units else v, round(v / total, 3)) for k, v in per_file.most_common()])
    for k, v in per_line.most_common(top):
        print("  %-28s %12s  %.4f%s" % (k, ("%.1f" % (v / units)) if units else v, v / total, source_text(k)))


if __name__ == "__main__":
    main()
