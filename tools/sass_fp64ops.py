"""Developer tool: in the densest fp64 window of a SASS dump, count fp64 instructions by the number of register
source operands that are NOT served by the operand reuse cache.  usage: python tools/sass_fp64ops.py dump.sass [win]"""
import re, sys, collections, bisect
ins = []
for l in open(sys.argv[1]):
    m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m: ins.append(m.group(2).strip())
W = int(sys.argv[2]) if len(sys.argv) > 2 else 480
idx = [i for i, t in enumerate(ins) if re.match(r'(@!?U?P\d+\s+)?D(FMA|MUL|ADD)', t)]
lo, bestc = 0, 0
for s in range(0, max(1, len(ins) - W)):
    c = bisect.bisect_left(idx, s + W) - bisect.bisect_left(idx, s)
    if c > bestc: bestc, lo = c, s
cnt = collections.Counter(); prevflags = {}; other = collections.Counter()
for t in ins[lo:lo + W]:
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*)', t)
    op = m.group(2); ops = [o.strip() for o in m.group(3).split(',')]
    b = op.split('.')[0]
    if b in ('DFMA', 'DMUL', 'DADD'):
        nreg = 0; newflags = {}
        for slot, o in enumerate(ops[1:]):
            mm = re.fullmatch(r'-?\|?R(\d+)(\.reuse)?\|?', o)
            if mm:
                r = int(mm.group(1))
                if prevflags.get(slot) != r: nreg += 1
                if mm.group(2): newflags[slot] = r
        cnt[(b, nreg)] += 1
        prevflags = newflags
    else:
        other[b] += 1
        prevflags = {}
print(f"window [{lo},{lo+W}) fp64={bestc}")
print(sorted(cnt.items()))
print(sorted(other.items(), key=lambda x: -x[1]))
