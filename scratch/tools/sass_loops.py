#!/usr/bin/env python
"""Static SASS metric for the tracking kernels: for every innermost loop that contains 256-bit gathers
(LDG.E.ENL2.256), the number of instructions in the loop body per gather (= per edge point) and their mix.

  python scratch/tools/sass_loops.py <file.so|file.cubin> <kernel name substring>
"""
import collections
import re
import subprocess
import sys


def functions(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;\s*/\*", line)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2)))
    return out


def classify(ins):
    ins = re.sub(r"^@!?U?P\d\s+", "", ins)
    op = ins.split()[0].split(".")[0]
    if op in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2"):
        return "fp32"
    if op in ("FSEL", "FSETP", "FMNMX", "SEL", "ISETP", "PLOP3", "LOP3", "PRMT", "SHF", "IADD3", "VIADD", "IMAD", "LEA", "MOV", "IABS", "UMOV", "USEL", "UISETP", "CS2R"):
        return "int/select"
    if op in ("I2F", "I2FP", "F2I", "MUFU", "F2F", "F2FP"):
        return "convert/mufu"
    if op in ("LDG", "LDS", "STS", "LDC", "LDCU", "LDL", "STL", "S2R", "S2UR"):
        return "memory/special"
    if op in ("BRA", "BSSY", "BSYNC", "WARPSYNC", "EXIT", "BAR"):
        return "control"
    return "other:" + op


def main():
    path, pat = sys.argv[1], sys.argv[2]
    for name, ins in functions(path).items():
        if pat not in name:
            continue
        addr_index = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, s) in enumerate(ins):
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", s)
            if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr_index:
                loops.append((addr_index[int(m.group(1), 16)], i))
        # innermost loops with gathers
        res = []
        for lo, hi in loops:
            body = ins[lo:hi + 1]
            g = sum("ENL2.256" in s for _, s in body)
            if g == 0:
                continue
            if any(lo2 >= lo and hi2 <= hi and (lo2, hi2) != (lo, hi) and any("ENL2.256" in s for _, s in ins[lo2:hi2 + 1]) for lo2, hi2 in loops):
                continue
            res.append((lo, hi, g, body))
        print(name[:90])
        for lo, hi, g, body in res:
            mix = collections.Counter(classify(s) for _, s in body)
            spills = sum(1 for _, s in body if re.search(r"\b(LDL|STL)\b", s))
            print(f"  loop @{ins[lo][0]:#x}..{ins[hi][0]:#x}: {len(body)} instructions, {g} gathers -> {len(body) / g:.1f} per point; "
                  f"spill ld/st {spills}; mix/point: " + ", ".join(f"{k} {v / g:.1f}" for k, v in sorted(mix.items())))


if __name__ == "__main__":
    main()
