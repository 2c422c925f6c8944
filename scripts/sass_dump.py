"""SASS evidence of the hot kernels (no GPU needed): opcode histogram + the Blackwell-specific mnemonics, per kernel.
    python scripts/sass_dump.py > profiles/<round>_sass.txt
NVRTC-specialised kernels are compiled here for sm_100a (engine.jit_check) and read back from the in-tree cubin cache; the offline kernels
come from libmcsolver_b200.so."""
import collections, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mcsolver_b200 import engine
from mcsolver_b200.lattice import add_dipole_stencil
from tests.specs import spec_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CACHE = os.path.join(ROOT, "mcsolver_b200", "build", "jitcache")
MARK = ["FFMA2", "FMUL2", "FADD2", "UBLKPF", "UBLKCP", "UTMALDG", "LDGSTS", "IDP", "POPC", "DFMA", "MUFU", "LDG.E.128", "STG.E.128", "LDL", "STL", "REDUX", "ATOMG", "RED"]


def hist(sass):
    ops = collections.Counter()
    marks = collections.Counter()
    n = 0
    for l in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if not m:
            continue
        toks = [t for t in m.group(1).split() if not t.startswith("@")]
        if not toks:
            continue
        n += 1
        ops[toks[0].split(".")[0]] += 1
        for k in MARK:
            if toks[0].startswith(k):
                marks[k] += 1
    return n, ops, marks


def report(title, sass, res=""):
    n, ops, marks = hist(sass)
    print("== %s\n   %d instructions  %s" % (title, n, res))
    print("   marks: " + "  ".join("%s=%d" % (k, v) for k, v in sorted(marks.items(), key=lambda kv: -kv[1])))
    print("   top:   " + "  ".join("%s=%d" % kv for kv in ops.most_common(14)))


def jit(title, spec, model, prec, fun="mcg_pass_m1", colour=1):
    n, rep = engine.jit_check(spec, model, prec)
    keys = re.findall(r"colour (\d+)(?: \(int8\))?: module ([0-9a-f]{16})", rep)
    topo = re.findall(r"topological charge: module ([0-9a-f]{16})", rep)
    key = topo[0] if fun == "mcg_topo" else dict(keys).get(str(colour), keys[-1][1])
    cub = os.path.join(CACHE, key + ".cubin")
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, cub], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", cub], capture_output=True, text=True).stdout
    m = re.search(r"Function %s:\s*\n\s*(.*)" % fun, res)
    report("%s  [%s, module %s]" % (title, fun, key), sass, m.group(1).strip() if m else "")


jit("headline: Heisenberg sc 256^3 fp32 colour pass with fused M,E", bench.cubic_spec(256), 3, 32)
jit("fp64 state: Heisenberg sc 256^3 colour pass (bulk L2 prefetch of the rows that miss)", bench.cubic_spec(256), 3, 64)
jit("C1: XY square 4096^2 fp32", bench.square_spec(4096), 2, 32)
jit("C2: Ising square 4096^2 int8 state", bench.square_spec(4096), 1, 8)
jit("C3: CrI3 honeycomb 512^2 x 2 fp32 (12 links, D)", spec_of("cri3", (512, 512, 1)), 3, 32)
jit("C4: skyrmion hex 1024^2 x 2 fp32 colour pass (DMI: full tensors)", spec_of("skyrmion", (1024, 1024, 1)), 3, 32)
jit("C4: topological charge", spec_of("skyrmion", (1024, 1024, 1)), 3, 32, fun="mcg_topo")
lib = os.path.join(ROOT, "mcsolver_b200", "libmcsolver_b200.so")
all_sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for pat, title in (("k_struct_async", "dipole stencil pass: asynchronous link pipeline (offline build, cp.async = LDGSTS)"),
                   ("k_wolff_frontier", "Wolff frontier growth"), ("k_wolff_bonds", "Wolff global bond pass")):
    blocks = re.split(r"\n\s*Function : ", all_sass)
    sel = [b for b in blocks if pat in b.split("\n", 1)[0]]
    if sel:
        b = max(sel, key=len)
        report("%s  [%s..., largest of %d instantiations]" % (title, b.split("\n", 1)[0][:60], len(sel)), b)
