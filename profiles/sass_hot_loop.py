"""Hot loop of a kernel from an object file: the longest straight-line SASS region by IMAD.WIDE count (sketch kernel) or, with
--pattern OPC, by that opcode; prints per-opcode and per-pipe instruction counts and the region itself.
usage: python profiles/sass_hot_loop.py <obj> <mangled-function-prefix> [windows-per-iteration] [--pattern LDS] [--listing]"""
import sys, re, subprocess, collections
obj, func = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
# split by function
parts = out.split("Function : ")
body = [p for p in parts if p.startswith(func)][0]
ins = []
for ln in body.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
# find the longest straight-line region containing IMAD.WIDE (hot loop): region between branch instrs with most IMAD.WIDE
regions, cur = [], []
for a, t in ins:
    cur.append((a, t))
    if re.search(r"\b(BRA|BSYNC|BSSY|EXIT|RET|CALL|WARPSYNC|BRA\.DIV)\b", t):
        regions.append(cur); cur = []
pat = sys.argv[sys.argv.index("--pattern") + 1] if "--pattern" in sys.argv else "IMAD.WIDE"
best = max(regions, key=lambda r: sum(pat in t for _, t in r))
print("hot region 0x%x..0x%x: %d instructions" % (best[0][0], best[-1][0], len(best)))
c = collections.Counter()
for _, t in best:
    op = t.split()[0]
    if op.startswith("@"): op = t.split()[1]
    c[op] += 1
args=[a for a in sys.argv[3:] if not a.startswith("--")]
win = int(args[0]) if args else 4
def pipe(op):
    b = op.split(".")[0]
    if b == "IMAD": return "fma_wide" if "WIDE" in op else "fma"
    if b in ("SHF","LOP3","IADD3","PRMT","ISETP","SEL","LEA","MOV","IADD","SGXT","BMSK","PLOP3","VIADD","FLO","POPC","VIMNMX"): return "alu"
    if b in ("LDS","STS","LDG","STG","LD","ST","ATOMS","ATOMG","RED"): return "lsu"
    return "other"
pc = collections.Counter()
for op, n in c.items(): pc[pipe(op)] += n
for op, n in sorted(c.items(), key=lambda x: -x[1]): print("  %-24s %4d  %6.2f/window" % (op, n, n / win))
print("per window:", {k: round(v / win, 2) for k, v in pc.items()}, "total %.1f" % (len(best) / win))
fma_clk = (pc["fma_wide"] * 4.1 + pc["fma"] * 2) / win
print("pipe clocks per warp-window: fma %.0f  alu %.0f  issue %.0f" % (fma_clk, pc["alu"] * 2 / win, len(best) / win))

if "--listing" in sys.argv:
    print()
    for a, t in best: print("    /*%04x*/  %s" % (a, t))
