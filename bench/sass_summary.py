"""Opcode histogram of the product kernels in the built library (read here, no GPU needed):
    python bench/sass_summary.py > profiles/r2_sass_summary.txt
Shows that the hot loops are packed fp32 (FFMA2 / FMUL2 with scalar-broadcast operands), that staging is bulk TMA
(UBLKCP) with mbarriers (SYNCS), that the fp64 kernels are DFMA/DMUL, and that no tensor-core or generic-pointer
load (LD) sits in them."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "gkl_b200" / "lib" / "libgkl_pairhmm.so"
WANT = ["k_h2_tasksILi8ELi13ELi8", "k_h2_tasksILi16ELi10ELi8", "k_h2_tasksILi4ELi16ELi8", "k_h2_megaILi8", "k_r2_listILi8ELi13ELi8",
        "k_sweep_tasksINS_3VF2ELi32ELi8ELi8ELb1", "k_sweep_listINS_3VD1ELi16ELi7", "k_sweep_tasksINS_3VD1ELi16ELi7",
        "k_mega_listINS_3VD1", "k_mega_tasksINS_3VD1", "k_pack_reads", "k_fill_pair_panel", "k_pdhmm2ILi4ELi12", "k_pdhmm2ILi5ELi8",
        "k_smith_waterman", "k_narrow", "k_collect_overrides_r2"]
OPS = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "DFMA", "DMUL", "DADD", "LDS", "STS", "LD", "LDG", "STG", "LDL", "STL",
       "SHFL", "UBLKCP", "SYNCS", "ATOMG", "HMMA", "UTCHMMA", "IMAD", "LOP3", "FSEL", "BRA", "CALL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s+Function : ", out)
    print(f"# {LIB.relative_to(ROOT)}: {len(funcs) - 1} device functions; arch lines: "
          f"{sorted(set(re.findall(r'arch = (sm_\w+)', out)))}")
    print("# scalar-broadcast FFMA2/FMUL2 operands (R.F32) in the whole library:", len(re.findall(r"(?:FFMA2|FMUL2)[^;]*R\d+\.F32,", out)))
    print(f"{'kernel':52s} " + " ".join(f"{o:>6s}" for o in OPS) + "   total")
    for f in funcs[1:]:
        name = f.split("\n")[0].strip()
        if not any(w in name for w in WANT):
            continue
        c = collections.Counter()
        n = 0
        for l in f.split("\n"):
            m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
            if m:
                c[m.group(1)] += 1
                n += 1
        short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", short).replace("gklb::", "").replace("void ", "")
        print(f"{short[:52]:52s} " + " ".join(f"{c.get(o, 0):6d}" for o in OPS) + f" {n:7d}")


if __name__ == "__main__":
    main()
