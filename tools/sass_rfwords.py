"""Developer tool: estimate register-file read words (32-bit) per instruction for a SASS loop body.
Usage: cuobjdump -sass lib.so | python tools/sass_rfwords.py <function-substring> [start_idx end_idx]
Counts source register operands (64-bit operands of D*/F2F.F64 sources, .64 etc. count 2 words), skipping operands
that hit the reuse cache (same register, same slot, flagged .reuse on the previous instruction)."""
import re, sys, collections
fn = sys.argv[1]
lines = sys.stdin.read().splitlines()
ins = []
on = False
for l in lines:
    if 'Function :' in l:
        on = fn in l
        continue
    if not on: continue
    m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
def parse(text):
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*)', text)
    op = m.group(2); ops = [o.strip() for o in m.group(3).split(',')] if m.group(3) else []
    return op, ops
WIDE_SRC = ('DFMA', 'DMUL', 'DADD', 'DSETP', 'F2F.F32.F64')
def words(op, ops, prev_reuse):
    base = op.split('.')[0]
    srcs = ops[1:] if base not in ('STS', 'STG', 'STL', 'BRA', 'ISETP', 'FSETP') else ops
    if base in ('ISETP', 'FSETP', 'DSETP'): srcs = ops[2:]
    w = 0; reuse_now = {}
    for slot, o in enumerate(srcs):
        mm = re.search(r'\bR(\d+)(\.reuse)?', o)
        if not mm or 'UR' in o.split('R' + mm.group(1))[0][-1:]:
            continue
        if re.search(r'\bUR\d+', o) and not re.search(r'(?<!U)R\d+', o): continue
        reg = int(mm.group(1))
        wide = 2 if (op.startswith(WIDE_SRC) or (base in ('LDS', 'LDG', 'STS') and False)) else 1
        if '[' in o: wide = 1
        if prev_reuse.get(slot) == reg: pass
        else: w += wide
        if mm.group(2): reuse_now[slot] = reg
    return w, reuse_now
a, b = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, len(ins))
tot = 0; cnt = collections.Counter(); wc = collections.Counter(); prev = {}
for addr, text in ins[a:b]:
    op, ops = parse(text)
    w, prev = words(op, ops, prev)
    tot += w; cnt[op.split('.')[0]] += 1; wc[op.split('.')[0]] += w
n = b - a
print(f"instructions {n}, rf words {tot}")
for k, v in sorted(wc.items(), key=lambda x: -x[1])[:20]:
    print(f"  {k:10s} n={cnt[k]:4d} words={v:5d}  ({v/cnt[k]:.2f}/instr)")
