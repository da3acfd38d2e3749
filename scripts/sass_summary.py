"""profiles/sass/*: per-kernel SASS statistics of the shipped library.  python scripts/sass_summary.py > profiles/sass/r02_sass_summary.txt"""
import collections, re, subprocess
lib = "katana.jl_b200/libktn.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
dem = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
for m, d in zip(re.findall(r"Function : (\S+)", sass), dem): names[m] = d
print("# cuobjdump -sass katana.jl_b200/libktn.so, per kernel: instruction count, local-memory instructions (STL / LDL = spills),")
print("# global loads (LDG; 256-bit = LDG.E.256, one per lane and group of the family blobs), bulk copies (UBLKCP = cp.async.bulk), fp64 arithmetic")
out = []
for blk in sass.split("Function : ")[1:]:
    mangled = blk.split()[0]
    ops = collections.Counter()
    n = 0
    for line in blk.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m: continue
        n += 1; op = m.group(1); ops[op.split(".")[0]] += 1
        if op.startswith("LDG") and ".256" in op: ops["LDG256"] += 1
    out.append(f"{names.get(mangled, mangled)[:110]:110s} instr {n:6d}  STL {ops['STL']:4d}  LDL {ops['LDL']:4d}  LDG {ops['LDG']:4d} (256-bit {ops['LDG256']:3d})  LDGSTS {ops['LDGSTS']:3d}  UBLKCP {ops['UBLKCP']:2d}  DFMA {ops['DFMA']:5d}  DMUL {ops['DMUL']:4d}  DADD {ops['DADD']:4d}")
print("\n".join(sorted(out)))
