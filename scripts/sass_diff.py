"""Compares the SASS of every kernel in two object files (cuobjdump -sass, instruction text only, addresses and encodings dropped).
usage: python scripts/sass_diff.py old.o new.o -- used to show that a refactoring left the GPU-verified kernels untouched."""
import subprocess, re, sys, hashlib
def funcs(obj):
    out = subprocess.check_output(["cuobjdump", "-sass", obj]).decode()
    res = {}; name = None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+_cu_[0-9a-f]+", "ANON", m.group(1)); res[name] = []; continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
        if m and name: res[name].append(m.group(1))
    return {k: hashlib.md5("\n".join(v).encode()).hexdigest() + ":%d" % len(v) for k, v in res.items()}
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
same = [k for k in a if k in b and a[k] == b[k]]
diff = [k for k in a if k in b and a[k] != b[k]]
print("same", len(same), "different", len(diff), "only old", [k for k in a if k not in b], "only new", [k for k in b if k not in a])
for k in diff: print("DIFF", k, a[k], b[k])
