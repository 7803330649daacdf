"""profiles/sass_<tag>.md: per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass) and short excerpts that
show the Blackwell-specific instructions in use.   usage: python tools/sass_summary.py TAG"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gypsum_b200", "libgypsum_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs, cur = {}, None
arch = set(re.findall(r"arch = (sm_\w+)", txt))
for line in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).rstrip())
demangled = dict(zip(funcs, subprocess.run(["c++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()))
WATCH = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "UBLKCP", "SYNCS", "LDGSTS", "REDUX", "LDS", "STS", "LDG", "STG",
         "BAR", "DFMA", "HMMA", "UTC"]


def op_of(line):
    body = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", line)
    body = re.sub(r"^@!?U?P\d+\s+", "", body)
    return body.split()[0].rstrip(";") if body.split() else ""


out = [f"# SASS of the shipped library ({tag})", "",
       f"`cuobjdump -sass gypsum_b200/libgypsum_b200.so`, cubin architectures: {', '.join(sorted(arch))}.  Counts are static instructions per kernel.",
       "`FFMA2 / FADD2 / FMUL2` = packed FP32 (sm_100), `UBLKCP` = TMA bulk copy (`cp.async.bulk`), `SYNCS` = mbarrier operations, `LDGSTS` = `cp.async`,",
       "`REDUX` = warp reduce; no tensor-core (`HMMA` / `UTC*MMA`) instruction anywhere, as the tier intends.", "",
       "| kernel | instr | " + " | ".join(WATCH) + " |", "|---|---|" + "---|" * len(WATCH)]
rows = []
for f, lines in funcs.items():
    c = collections.Counter()
    for ln in lines:
        op = op_of(ln)
        base = op.split(".")[0]
        for w in WATCH:
            if base == w or (w == "UTC" and base.startswith("UTC")):
                c[w] += 1
    name = re.sub(r"\(.*", "", demangled.get(f, f)).replace("void gb::", "")
    rows.append((name, len(lines), c))
for name, n, c in sorted(rows, key=lambda r: -r[1]):
    out.append(f"| `{name}` | {n} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |")


def excerpt(func_substr, pattern, before=2, after=6, title=""):
    for f, lines in funcs.items():
        if func_substr in demangled.get(f, f):
            for i, ln in enumerate(lines):
                if re.search(pattern, ln):
                    out.extend(["", f"### {title}", "", "```"] + lines[max(0, i - before): i + after] + ["```"])
                    return


excerpt("k_correlate_w2048<12, true>", r"UBLKCP", 6, 4, "k_correlate_w2048<12,true>: replica spectrum staged by a TMA bulk copy behind an mbarrier")
excerpt("k_correlate_w2048<12, true>", r"FFMA2 .*LO_HI", 3, 9, "k_correlate_w2048<12,true>: packed butterflies (half swap + per-lane sign on the operand, splat immediates)")
excerpt("k_correlate_w2048<12, true>", r"REDUX", 2, 6, "k_correlate_w2048<12,true>: peak / first index / count merged with REDUX")
excerpt("k_track_channels<2>", r"LDGSTS", 2, 4, "k_track_channels<2>: next millisecond prefetched with cp.async (LDGSTS) while the current one is processed")
excerpt("k_acquire_fused<2, 2>", r"UBLKCP", 4, 4, "k_acquire_fused<2,2>: IQ chunk and replica spectrum by TMA bulk copies")
open(os.path.join(ROOT, "profiles", f"sass_{tag}.md"), "w").write("\n".join(out) + "\n")
print("wrote profiles/sass_%s.md" % tag, len(rows), "kernels")
