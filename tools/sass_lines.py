"""SASS instruction count per source line for one device function of libdirect_ddp_b200.so (needs -lineinfo).
    python tools/sass_lines.py rows_unitId [lib.so]      # substring of the mangled name
"""
import collections, os, re, subprocess, sys, tempfile
pat = sys.argv[1]
lib = os.path.abspath(sys.argv[2] if len(sys.argv) > 2 else "direct_b200/libdirect_ddp_b200.so")
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cub)], capture_output=True, text=True).stdout
cur = None; func = None; on = False
cnt = collections.Counter(); tot = collections.Counter()
for line in txt.split("\n"):
    m = re.match(r'^(\$?[_\w\$]+):\s*$', line)
    if m:
        func = m.group(1); on = pat in func and not func.startswith(".L")
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if on and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
        cnt[cur] += 1; tot[func] += 1
for f, v in tot.items():
    print(v, "SASS instructions in", f[-70:])
src = {}
for (f, l), v in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
    if f not in src:
        try: src[f] = open(os.path.join("direct_b200/csrc", f)).read().split("\n")
        except OSError: src[f] = []
    t = src[f][l - 1].strip()[:90] if l - 1 < len(src[f]) else ""
    print(f"  {v:5d}  {f}:{l:<5d} {t}")
