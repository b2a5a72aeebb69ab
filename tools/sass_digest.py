"""Per-kernel instruction-class counts of weiner_slamit_v2_b200/liborb_b200.so (cuobjdump -sass), so the instruction-mix claims
in DESIGN.md can be checked without rebuilding: regenerate with every kernel change.

    python tools/sass_digest.py > profiles/r2_sass_digest.json
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "weiner_slamit_v2_b200", "liborb_b200.so")
CLASSES = [  # (name, regex on the mnemonic with its suffixes)
    ("LDG.E.128", r"^LDG\.E\.(?:\w+\.)*128"), ("LDG.E.64", r"^LDG\.E\.(?:\w+\.)*64"), ("LDG.E.U8/U16", r"^LDG\.E\.(?:\w+\.)*(?:U8|S8|U16|S16)"),
    ("LDG.E.32", r"^LDG\.E(?!\.(?:\w+\.)*(?:128|64|U8|S8|U16|S16))"),
    ("STG", r"^STG"), ("UBLKCP (TMA 1-D bulk)", r"^UBLKCP"), ("UTMALDG (TMA tensor)", r"^UTMALDG"), ("UTMASTG", r"^UTMASTG"),
    ("LDS.128", r"^LDS\.(?:\w+\.)*128"), ("LDS.64", r"^LDS\.(?:\w+\.)*64"), ("LDS.U8/U16", r"^LDS\.(?:\w+\.)*(?:U8|S8|U16|S16)"),
    ("LDS.32", r"^LDS(?!\.(?:\w+\.)*(?:128|64|U8|S8|U16|S16))"),
    ("STS", r"^STS"), ("ATOMS/ATOMG/RED", r"^(?:ATOMS|ATOMG|ATOM|RED)"),
    ("IDP (dp4a/dp2a)", r"^IDP"), ("VIMNMX3", r"^VIMNMX3"), ("VIMNMX", r"^VIMNMX(?!3)"), ("VIADDMNMX", r"^VIADDMNMX"), ("VABSDIFF", r"^VABSDIFF"),
    ("POPC", r"^POPC"), ("PRMT", r"^PRMT"), ("LOP3", r"^LOP3"), ("IADD3/VIADD", r"^(?:IADD3|VIADD|IADD)"), ("IMAD", r"^IMAD"), ("SHF/SHL/SHR", r"^(?:SHF|SHL|SHR)"),
    ("ISETP/SEL", r"^(?:ISETP|SEL|FSEL|FSETP|PLOP3)"), ("VOTE/BALLOT", r"^VOTE"), ("SHFL", r"^SHFL"), ("REDUX", r"^REDUX"), ("MATCH", r"^MATCH"),
    ("FP32 (FADD/FMUL/FFMA/FMNMX)", r"^(?:FADD|FMUL|FFMA|FMNMX)"), ("FP64", r"^(?:DADD|DMUL|DFMA|DSETP)"), ("MUFU", r"^MUFU"),
    ("conversions (I2F/F2I/F2F/I2I)", r"^(?:I2F|F2I|F2F|I2I|I2FP|F2FP)"), ("BAR/SYNCS (mbarrier)", r"^(?:BAR|SYNCS|WARPSYNC)"),
    ("BRA/EXIT/CALL", r"^(?:BRA|EXIT|CALL|RET|BSSY|BSYNC|BREAK)"),
    ("LEA", r"^LEA"), ("MOV/CS2R/S2R", r"^(?:MOV|CS2R|S2R|S2UR|R2UR|UMOV)"), ("LDC/ULDC (constant bank)", r"^(?:LDC|ULDC|LDCU)"),
    ("uniform datapath (U*)", r"^U[A-Z]"), ("NOP", r"^NOP"), ("LDGSTS/LDGDEPBAR", r"^(?:LDGSTS|LDGDEPBAR|DEPBAR)"), ("FENCE/MEMBAR/CCTL", r"^(?:FENCE|MEMBAR|CCTL|ERRBAR)"),
]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else SO
    txt = subprocess.check_output(["cuobjdump", "-sass", so], universal_newlines=True)
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.check_output(["c++filt", m.group(1)], universal_newlines=True).strip()
            name = re.sub(r"\(.*\)$", "", name).replace("orbb200::", "").replace("(anonymous namespace)::", "")
            name = re.sub(r"^void ", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for cname, rx in CLASSES:
                if re.match(rx, op):
                    cur[cname] += 1
                    break
            else:
                cur["other"] += 1
    out = {"library": os.path.relpath(so, ROOT), "how": "cuobjdump -sass, static instruction counts per kernel (not executed counts)",
           "kernels": {k: dict(sorted(v.items(), key=lambda kv: -kv[1])) for k, v in kernels.items()}}
    json.dump(out, sys.stdout, indent=1)
    sys.stdout.write("\n")


if __name__ == "__main__":
    main()
