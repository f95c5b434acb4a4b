"""Offline look at what NVRTC/ptxas make of a program variant (no GPU needed): registers, spills, SASS size and the
issue-stall counts of the control codes.  python scripts/sass_report.py {lorenz|rober|pleiades} ALG {f64|f32} [extra options] [--dump file]"""
import os, re, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200_import
pkg = b200_import.load()
pl = pkg.problems_library


def decode(cubin_path, fn="b200_integrate"):
    txt = subprocess.run(["cuobjdump", "-sass", cubin_path], stdout=subprocess.PIPE, text=True).stdout.split("\n")
    out, infn, i = [], False, 0
    while i < len(txt):
        l = txt[i]
        m = re.search(r"Function : (\S+)", l)
        if m:
            infn = (m.group(1) == fn); i += 1; continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", l)
        if m and infn:
            hi = int(re.search(r"/\* (0x[0-9a-f]+) \*/", txt[i + 1]).group(1), 16)
            out.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xf, (hi >> 45) & 1, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3f))
            i += 2; continue
        i += 1
    return out


def main():
    prob, alg, ty = sys.argv[1], sys.argv[2], sys.argv[3]
    rest = sys.argv[4:]
    dump = None
    if "--dump" in rest:
        k = rest.index("--dump"); dump = rest[k + 1]; rest = rest[:k] + rest[k + 2:]
    extra = " ".join(rest) or None
    f32 = ty == "f32"
    algid = getattr(pkg, "ALG_" + alg.upper())
    jac = tg = None
    if prob == "lorenz":
        rhs = pl.lorenz_source(f32); n, np_ = 3, 3
    elif prob == "rober":
        rhs, jac, tg = pl.robertson_sources(f32); n, np_ = 3, 3
    else:
        rhs = pl.pleiades_source(f32, loops=True); n, np_ = 28, 0
    cubin, log = pkg.compile_only(algid, pkg.F32 if f32 else pkg.F64, n, np_, rhs[0], rhs[1], jac[0] if jac else None,
                                  jac[1] if jac else None, tg[0] if tg else None, tg[1] if tg else None, extra_options=extra)
    for l in log.split("\n"):
        if "registers" in l or "spill" in l or "Compiling entry" in l:
            print(l.strip())
    with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
        f.write(cubin); path = f.name
    ins = decode(path)
    os.unlink(path)
    from collections import Counter
    c = Counter(i[1].split()[1 if i[1].startswith("@") else 0].split(".")[0] for i in ins)
    print("b200_integrate: %d SASS instructions; static stall sum %d" % (len(ins), sum(i[2] for i in ins)))
    print("  " + " ".join("%s:%d" % kv for kv in c.most_common(16)))
    if dump:
        with open(dump, "w") as f:
            for a, s, st, y, wb, rb, wm in ins:
                f.write("%05x s=%2d y=%d wb=%s rb=%s wait=%02x  %s\n" % (a, st, y, wb if wb != 7 else "-", rb if rb != 7 else "-", wm, s))


main()
