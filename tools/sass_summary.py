#!/usr/bin/env python
"""Static view of the built library (no GPU needed): per kernel, the resource usage ptxas settled on and how often
the SASS mnemonics that identify the Blackwell paths occur (B200_PROFILING.md, "What proves a Blackwell-native kernel").

    python tools/sass_summary.py > profiles/rNN_sass_static.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hyperparameter-gnn_unfolded-d-admm-main_b200", "libdadmm_sm100.so")
MNEMONICS = [("UTC*MMA", r"\bUTC[A-Z]*MMA\b", "tcgen05.mma"), ("LDTM", r"\bLDTM\b", "tcgen05.ld"),
             ("UTMALDG", r"\bUTMALDG\b", "TMA tensor load"), ("UBLKCP", r"\bUBLKCP\b", "TMA bulk copy"),
             ("LDGSTS", r"\bLDGSTS\b", "cp.async"), ("SYNCS", r"\bSYNCS\b", "mbarrier"),
             ("ACQBULK", r"\bACQBULK\b", "griddepcontrol.wait"), ("PREEXIT", r"\bPREEXIT\b", "griddepcontrol.launch_dependents"),
             ("HMMA", r"\bHMMA\b", "legacy mma.sync (none expected)"), ("STL/LDL", r"\b(STL|LDL)\b", "local-memory spills"),
             ("F*2", r"\b(FADD2|FMUL2|FFMA2)\b", "packed fp32x2 arithmetic"), ("SHFL", r"\bSHFL\b", "warp shuffles"), ("FFMA", r"\bFFMA\b", "fp32 FMA"), ("DFMA", r"\bDFMA\b", "fp64 FMA")]


def demangle(names):
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)          # argument list
    return name.replace("dadmm::", "").replace("(anonymous namespace)::", "")


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    for fn, line in re.findall(r"Function ([^:\n]+):\n\s*(REG:[^\n]+)", res):
        usage[fn] = dict(kv.split(":") for kv in line.split())
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, size, cur = collections.defaultdict(collections.Counter), collections.Counter(), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        if cur is None or not re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            continue
        size[cur] += 1
        for key, pat, _ in MNEMONICS:
            if re.search(pat, ln):
                counts[cur][key] += 1
    names = demangle(sorted(usage))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(usage)} kernels, sm_100a; columns = registers / static+dynamic-independent shared bytes /")
    print("# local-memory stack bytes / SASS instructions, then occurrences of the identifying mnemonics")
    print("# " + "; ".join(f"{k} = {what}" for k, _, what in MNEMONICS))
    rows = []
    for fn in sorted(usage, key=lambda f: short(names[f])):
        u, c = usage[fn], counts[fn]
        marks = " ".join(f"{k}={c[k]}" for k, _, _ in MNEMONICS if c[k])
        rows.append(f"{short(names[fn])[:110]:<110} REG={u.get('REG', '?'):>3} SHARED={u.get('SHARED', '?'):>6} STACK={u.get('STACK', '?'):>4} "
                    f"INSTR={size[fn]:>5}  {marks}")
    print("\n".join(rows))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("# totals: " + " ".join(f"{k}={tot[k]}" for k, _, _ in MNEMONICS))
    spills = [short(names[f]) for f in usage if int(usage[f].get("STACK", 0)) > 0]
    print(f"# kernels with a local-memory stack frame: {len(spills)}" + (": " + "; ".join(s[:60] for s in spills) if spills else ""))


if __name__ == "__main__":
    sys.exit(main())
