#!/usr/bin/env python3
"""Join an `ncu --page source --csv` SASS listing with `nvdisasm --print-line-info` of the same cubin
and print executed warp-instructions and stall samples per source line.

usage: ncu_by_line.py <source_page.csv> <nvdisasm_lines.txt> <mangled kernel name> [top N]
"""
import csv, re, sys, collections

def main():
    src_csv, dis, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = [r for r in csv.reader(open(src_csv)) if r and r[0].startswith("0x")]
    hdr = next(r for r in csv.reader(open(src_csv)) if r and r[0] == "Address")
    ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
    lines, cur, on = [], ("?", 0), False
    for ln in open(dis):
        if ln.startswith(".text."):
            on = ln.strip() == ".text.%s:" % kern
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((cur, m.group(2).strip()))
    n = len(lines)
    if len(rows) % n:
        sys.exit("instruction count mismatch: csv %d, disasm %d" % (len(rows), n))
    rows = rows[:n]
    agg = collections.Counter(); smp = collections.Counter()
    tot = 0
    for (loc, txt), r in zip(lines, rows):
        op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
        if op.split(".")[0] not in txt:
            sys.exit("opcode mismatch at %s: %s vs %s" % (r[0], r[1], txt))
        agg[loc] += int(r[ie]); smp[loc] += int(r[ss]); tot += int(r[ie])
    st = sum(smp.values())
    print("total warp-instructions %d, samples %d" % (tot, st))
    for loc, v in agg.most_common(top):
        print("%-18s %5d  %6.2f%% inst  %6.2f%% samples" % (loc[0], loc[1], 100.0 * v / tot, 100.0 * smp[loc] / max(st, 1)))

main()
