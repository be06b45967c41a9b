"""Static SASS breakdown of one kernel: instructions between CTA barriers, integer / float / memory classes per segment.
Offline companion of scripts/ncu_segments.py (which needs an ncu source page): run after a build to see what a source
change did to the instruction stream before spending GPU time.

  python scripts/sass_segments.py raw2logit_b200/csrc/_obj/isp_bwd5_f32.o 'Bwd5CfgILi32ELi64ELi256ELb0ELb0EEEfLi2'
"""
import collections
import re
import subprocess
import sys

INT = ("IADD", "IMAD", "LEA", "SHF", "LOP", "ISETP", "MOV", "SEL", "IABS", "IMNMX", "PRMT", "R2UR", "S2R", "CS2R", "VIADD",
       "UMOV", "ULOP", "UIADD", "USHF", "ULEA", "UIMAD", "PLOP", "P2R", "R2P", "I2F", "F2I", "S2UR", "USEL", "UISETP", "VIMNMX")
FP = ("FFMA", "FMUL", "FADD", "FSEL", "FMNMX", "MUFU", "FSET", "FSETP", "DADD", "DFMA", "DMUL", "F2F", "FCHK")
MEM = ("LDG", "STG", "LDS", "STS", "LDC", "LDTM", "STTM", "ATOM", "RED", "LDL", "STL", "UTMA", "UBLK", "SYNCS", "ULDC")


def cls(op):
    for name, group in (("int", INT), ("fp", FP), ("mem", MEM)):
        if any(op.startswith(g) for g in group):
            return name
    return "ctl"


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, rows = None, []
    for line in names.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn and pat in fn:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                txt = m.group(2).strip()
                toks = txt.split()
                op = toks[1] if toks[0].startswith("@") else toks[0]
                rows.append((int(m.group(1), 16), op, txt))
    if not rows:
        sys.exit(f"no function matching {pat}")
    seg, segs = [], []
    for r in rows:
        seg.append(r)
        if r[1].startswith("BAR"):
            segs.append(seg)
            seg = []
    segs.append(seg)
    print(f"{len(rows)} instructions, {len(segs)} segments")
    for i, s in enumerate(segs):
        c = collections.Counter(cls(op) for _, op, _ in s)
        ops = collections.Counter(op.split(".")[0] for _, op, _ in s)
        top = " ".join(f"{k}:{v}" for k, v in ops.most_common(8))
        print(f"seg{i:2d} {s[0][0]:#07x} n {len(s):5d} int {c['int']:4d} fp {c['fp']:4d} mem {c['mem']:4d} ctl {c['ctl']:4d} | {top}")


if __name__ == "__main__":
    main()
