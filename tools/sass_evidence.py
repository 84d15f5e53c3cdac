"""SASS evidence table of libbem_b200.so (no GPU needed): per kernel, summed over its template instantiations, the counts of the
mnemonics B200_PROFILING.md names as proof of a Blackwell-native kernel.  python tools/sass_evidence.py > profiles/rNN_sass_evidence.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "bayesian-enhancement-model_b200", "libbem_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "FFMA2", "FMUL2", "FADD2", "HMMA", "MUFU"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = re.sub(r"<.*", "", demangle(m.group(1)))
        name = re.sub(r"^void ", "", name)
        cur = per.setdefault(name, dict(inst=0, n=0, ops=collections.Counter()))
        cur["n"] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur["inst"] += 1
        op = m.group(1)
        if op in WATCH:
            cur["ops"][op] += 1
regs = collections.defaultdict(list)
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", res):
    name = re.sub(r"^void ", "", re.sub(r"<.*", "", demangle(m.group(1))))
    regs[name].append((int(m.group(2)), int(m.group(3))))
print("| kernel | instantiations | SASS instructions | registers (min-max) | stack B (max) | tcgen05 / TMEM | TMA / async copies | mbarrier | packed fp32 (FFMA2 / FMUL2 / FADD2) | MUFU |")
print("|---|---:|---:|---|---:|---|---|---:|---|---:|")
tot_h = 0
for name, d in per.items():
    o = d["ops"]
    tot_h += o["HMMA"]
    tc = ", ".join(f"{k} {o[k]}" for k in ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS") if o[k]) or "–"
    tma = ", ".join(f"{k} {o[k]}" for k in ("UBLKCP", "UTMALDG", "LDGSTS") if o[k]) or "–"
    pk = " / ".join(str(o[k]) for k in ("FFMA2", "FMUL2", "FADD2")) if (o["FFMA2"] or o["FMUL2"] or o["FADD2"]) else "–"
    r = regs.get(name, [])
    rr = f"{min(x[0] for x in r)}-{max(x[0] for x in r)}" if r else ""
    st = max((x[1] for x in r), default=0)
    print(f"| `{name}` | {d['n']} | {d['inst']} | {rr} | {st} | {tc} | {tma} | {o['SYNCS'] or '–'} | {pk} | {o['MUFU'] or '–'} |")
print(f"\n`HMMA` (legacy `mma.sync`) occurrences in the whole library: {tot_h}.")
