"""Attribute the executed warp-instructions of one kernel in an .ncu-rep to SOURCE lines (innermost inlined frame):
share of instructions, active lanes per instruction and divergent branches per line.  ncu's CSV source page is per SASS
address; the line table comes from `nvdisasm -gi` of the cubin inside the built library (needs -lineinfo).

    cuobjdump -xelf all dexterous_rl_manipulation_b200/libdexsim_b200.so && nvdisasm -gi dexsim_kernels.sm_100a.cubin > all.sass
    python tools/ncu_lines.py report.ncu-rep <mangled kernel name prefix> all.sass [top_n]
"""
import csv, io, re, subprocess, sys, collections
rep, func_prefix, sass_file = sys.argv[1], sys.argv[2], sys.argv[3]
lines = open(sass_file).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + func_prefix))
addr2 = {}; chain = []; last_was_annot = False
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if not last_was_annot: chain = []
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        if m.group(3): chain.append((m.group(3).split("/")[-1], int(m.group(4))))
        last_was_annot = True
        continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(.*)', l)
    if m:
        last_was_annot = False
        if chain: addr2[int(m.group(1), 16)] = (chain[0], chain[-1], m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
col = {h: i for i, h in enumerate(hdr)}
ie = col["Instructions Executed"]; te = col.get("Thread Instructions Executed"); db = col.get("Divergent Branches")
leaf = collections.Counter(); leaf_thr = collections.Counter(); leaf_div = collections.Counter(); total = 0; seen = 0; base = None
for r in rows:
    if r and r[0] == "Kernel Name":
        seen += 1
        if seen > 1: break
        continue
    if len(r) < len(hdr) or r[0] == "Address": continue
    a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    if base is None: base = a
    n = int(r[ie] or 0); total += n
    ent = addr2.get(a - base)
    key = ent[0] if ent else ("?", 0)
    leaf[key] += n
    if te is not None: leaf_thr[key] += int(r[te] or 0)
    if db is not None: leaf_div[key] += int(r[db] or 0)
print("total warp-instr", total, "columns:", [h for h in hdr if "ranch" in h or "Thread" in h][:8])
for k, n in leaf.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 40):
    print(f"  {k[0]}:{k[1]:5d}  {100*n/total:5.1f}%  lanes/instr {leaf_thr[k]/max(n,1):5.1f}  divergent branches {leaf_div[k]}")
