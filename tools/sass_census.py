"""Instruction census of libqconsensus.so per kernel (cuobjdump -sass): the evidence that the hot path is hand-written
Blackwell code -- tcgen05 MMAs (UTCHMMA / UTCQMMA...), TMEM loads / stores (LDTM / STTM), TMA (UTMALDG, incl. GATHER4 / MULTICAST / 2CTA),
tcgen05.commit (UTCBAR), mbarrier traffic (SYNCS), legacy mma.sync (HMMA / IMMA) -- written as a markdown table.

  python tools/sass_census.py > profiles/sass_census_r2.md      (needs no GPU)"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "quoracle_b200" / "lib" / "libqconsensus.so"
COLS = OrderedDict([
    ("UTCHMMA", r"\bUTCHMMA"), ("…2CTA", r"\bUTCHMMA\S*\.2CTA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
    ("UTMALDG", r"\bUTMALDG"), ("…GATHER4", r"\bUTMALDG\S*GATHER4"), ("…MULTICAST", r"\bUTMALDG\S*MULTICAST"),
    ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"), ("HMMA (mma.sync)", r"\bHMMA"), ("MUFU.TANH", r"\bMUFU\.TANH"),
    ("ACQBULK/PDL", r"\bACQBULK|\bPREEXIT"),
])


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("qc::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"<EncTraits<([^>]*)>\s*>", lambda m: "<" + m.group(1).replace(" ", "") + ">", name)
    return name.replace(" ", "")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    funcs, cur = OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None and "/*" in line:
            funcs[cur].append(line)
    names = demangle(list(funcs))
    print("# SASS instruction census of `quoracle_b200/lib/libqconsensus.so` (sm_100a), round 2\n")
    print("`cuobjdump -sass`, counted per kernel by `tools/sass_census.py` (static instruction counts, not executed counts).")
    print("K1 = `qc_encode_kernel<dim,token block,cta group,W stage k,gather mode,multicast pairs>`; every instantiation is listed.\n")
    print("| kernel | SASS lines | " + " | ".join(COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    tot = [0] * len(COLS)
    for f, lines in funcs.items():
        text = "\n".join(lines)
        counts = [len(re.findall(rx, text)) for rx in COLS.values()]
        tot = [a + b for a, b in zip(tot, counts)]
        print(f"| `{short(names[f])}` | {len(lines)} | " + " | ".join(str(c) if c else "·" for c in counts) + " |")
    print(f"| **total ({len(funcs)} kernels)** | | " + " | ".join(f"**{c}**" for c in tot) + " |")
    lib_kernels = [f for f in funcs if re.search(r"cublas|cutlass|cudnn|triton", f, re.I)]
    print(f"\nLibrary kernels (cuBLAS / CUTLASS / cuDNN / Triton) in the binary: {len(lib_kernels)}.")


if __name__ == "__main__":
    sys.exit(main())
