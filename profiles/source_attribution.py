import re, csv, collections
lines = open('/tmp/p3_lines.txt', errors='replace').read().split('\n')
start = None
for i, l in enumerate(lines):
    if '.section' in l and 'text._ZN3agb23agb_newton_solve_kernelILi3ELi0ELi0E' in l:
        start = i; break
cur = None; off2line = {}; off2op = {}
for l in lines[start+1:]:
    if l.strip().startswith('.section') and 'text.' in l: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', l)
    if m and cur: off2line[int(m.group(1), 16)] = cur; off2op[int(m.group(1), 16)] = m.group(2).split('.')[0]
rows = list(csv.reader(open('/tmp/ncu_b_src.csv')))
hdr = rows[1]; ia = hdr.index("Address"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
base = int(rows[2][ia], 16)
print(len(rows)-2, "sass rows;", len(off2line), "mapped")
src = open('/root/repo/algames.jl_b200/csrc/agb_solver.cuh').read().split('\n')
fn_re = re.compile(r'__device__ .*?\b([a-zA-Z_][a-zA-Z0-9_]*)\s*\([^;]*\)\s*(const)?\s*\{')
marks = []
for i, s in enumerate(src):
    m = fn_re.search(s)
    if m and not s.strip().startswith('//'): marks.append((i+1, m.group(1)))
def region(f, ln):
    if f != 'agb_solver.cuh': return f
    name = 'other'
    for i, nm in marks:
        if i <= ln: name = nm
    return name
byline = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0; ops = collections.Counter()
for r in rows[2:]:
    off = int(r[ia], 16) - base
    n = int(r[ie]); s = int(r[isamp]); tot += n; tots += s
    key = off2line.get(off, ("?", 0))
    byline[key] += n; samp[key] += s; ops[off2op.get(off, '?')] += n
print("total warp instr", tot, "samples", tots)
reg = collections.Counter(); regs = collections.Counter()
for (f, ln), n in byline.items(): reg[region(f, ln)] += n
for (f, ln), n in samp.items(): regs[region(f, ln)] += n
for k, v in reg.most_common(22): print(f"{k:26s} instr {100*v/tot:5.1f}%  samples {100*regs[k]/tots:5.1f}%")
print("--- opcode mix (% of all instructions)")
print(", ".join(f"{k} {100*v/tot:.1f}" for k, v in ops.most_common(24)))
print("--- top lines")
for (f, ln), n in byline.most_common(28):
    if f == 'agb_solver.cuh': print(f"{100*n/tot:5.1f}% instr {100*samp[(f,ln)]/tots:5.1f}% samp  L{ln} [{region(f,ln)}]: {src[ln-1].strip()[:105]}")
    else: print(f"{100*n/tot:5.1f}% instr {100*samp[(f,ln)]/tots:5.1f}% samp  {f}:{ln}")
