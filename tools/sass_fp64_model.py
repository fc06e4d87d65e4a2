#!/usr/bin/env python
"""Estimate FP64-pipe issue cycles of a kernel's hot loop from its SASS.

Model measured with tools/dfma_probe.cu on B200: an FP64 instruction occupies the pipe
for max(2, R) cycles where R = number of distinct 64-bit REGISTER source operands that are
not served by the operand-reuse cache (uniform registers, immediates and constant-bank
operands cost no register-file read).

usage: sass_fp64_model.py <lib.so|cubin> <mangled-name-substring> [start_addr end_addr]
"""
import re
import subprocess
import sys


def main():
    lib, name = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    lines, on = [], False
    for ln in out.splitlines():
        if "Function :" in ln:
            on = name in ln
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", ln)
            if m:
                lines.append((int(m.group(1), 16), m.group(2).strip()))
    if len(sys.argv) > 4:
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    else:
        # innermost loop: last backward branch target .. branch
        lo = hi = None
        for addr, txt in lines:
            m = re.search(r"BRA(\.U)?\s+(UP\d+,\s*|!UP\d+,\s*)?0x([0-9a-f]+)", txt)
            if m and int(m.group(3), 16) < addr:
                span = addr - int(m.group(3), 16)
                if lo is None or span > hi - lo:
                    pass
                # choose the backward branch with the most FP64 instructions inside
                cnt = sum(1 for a, t in lines if int(m.group(3), 16) <= a <= addr and re.match(r"(@\S+\s+)?D(FMA|MUL|ADD)", t))
                if lo is None or cnt > best:
                    lo, hi, best = int(m.group(3), 16), addr, cnt
    reuse = {}  # slot -> register kept in the reuse cache
    tot = {"n": 0, "cyc": 0, "r3": 0, "mufu": 0, "other": 0}
    for addr, txt in lines:
        if not (lo <= addr <= hi):
            continue
        t = re.sub(r"^@\S+\s+", "", txt)
        m = re.match(r"(DFMA|DMUL|DADD)\S*\s+(.*)", t)
        if not m:
            if t.startswith("MUFU"):
                tot["mufu"] += 1
            else:
                tot["other"] += 1
            continue
        ops = [o.strip() for o in m.group(2).split(",")][1:]
        regs = set()
        new_reuse = {}
        for slot, o in enumerate(ops):
            mm = re.match(r"[-|]?(R\d+)(\.reuse)?", o)
            if not mm or o.startswith(("UR", "-UR", "c[", "-c[")):
                continue
            r = mm.group(1)
            if reuse.get(slot) != r:
                regs.add(r)
            if mm.group(2):
                new_reuse[slot] = r
        reuse = new_reuse
        c = max(2, len(regs))
        tot["n"] += 1
        tot["cyc"] += c
        tot["r3"] += c >= 3
    print(f"loop 0x{lo:04x}..0x{hi:04x}: {tot['n']} FP64 instr, {tot['cyc']} modelled pipe cycles "
          f"({tot['r3']} three-register), {tot['mufu']} MUFU, {tot['other']} other")


if __name__ == "__main__":
    main()
