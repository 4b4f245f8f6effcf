"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the built extension (runs on the CPU box).
usage: python tools/sass_summary.py > profiles/sass_tensor_tma_summary_r1.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "bert_pytorch_b200", "ops", "_C.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCMMA", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "LDTM", "STTM", "SYNCS", "LDGMC",
       "STGMC", "REDGMC", "MULTIMEM", "REDG", "ATOMG", "UTCATOM", "UTMACCTL", "UTMACMDFLUSH", "LDG.E.128.STRONG.SYS",
       "STG.E.128.STRONG.SYS", "LDG.E.64.STRONG.SYS"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    print("# cuobjdump -sass bert_pytorch_b200/ops/_C.so (sm_100a): per kernel, counts of tcgen05 / TMA / TMEM / mbarrier /\n"
          "# multimem SASS mnemonics.  UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = kind::f8f6f4, UTMALDG/UTMASTG = TMA load/store,\n"
          "# LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, SYNCS = mbarrier, .2CTA = cta_group::2 forms,\n"
          "# LDGMC = multimem.ld_reduce over NVLS, *.STRONG.SYS = system-scope peer loads/stores, REDG = red.global (incl. peer push)\n")
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        ins = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]+)", f)
        c = collections.Counter()
        for i in ins:
            for p in PAT:
                if i.startswith(p):
                    c[i] += 1
                    break
        if c:
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            dem = re.sub(r"\(.*", "", dem)
            print(f"{dem}  [{len(ins)} SASS instrs]\n    " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))


if __name__ == "__main__":
    main()
