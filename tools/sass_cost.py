#!/usr/bin/env python
"""Static issue-cost estimate of the loops of a kernel from its SASS (no GPU needed).

    python tools/sass_cost.py <file.cubin|file.so> <kernel-name regex> [--dump]

For every backward branch (loop) of the matching kernels prints the number of instructions of the loop body, the sum of the
stall counts ptxas encoded in the control words (the minimum number of issue cycles of one trip for a warp that owns its
scheduler: fixed-latency dependencies are resolved by these counts, variable-latency ones by scoreboard waits on top) and an
opcode histogram.  Control word layout (sm_70+, 128-bit instructions): bits 105-108 stall, 109 yield, 110-112 write barrier,
113-115 read barrier, 116-121 wait mask, 122-125 reuse.
"""
import collections
import re
import subprocess
import sys


def disassemble(path):
    out = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels, cur = {}, None
    lines = out.splitlines()
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = []
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", ln)
        if m and cur is not None:
            addr, text, lo = int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16)
            hi = 0
            if i + 1 < len(lines):
                m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
                if m2:
                    hi = int(m2.group(1), 16)
                    i += 1
            kernels[cur].append((addr, text, hi))
        i += 1
    return kernels


def ctrl(hi):
    return {"stall": (hi >> 41) & 0xF, "yield": (hi >> 45) & 1, "wbar": (hi >> 46) & 7, "rbar": (hi >> 49) & 7, "wait": (hi >> 52) & 0x3F}


def opcode(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    return t.split()[0].split(".")[0] if t else "?"


def main():
    path, pat = sys.argv[1], re.compile(sys.argv[2])
    dump = "--dump" in sys.argv
    for name, ins in disassemble(path).items():
        if not pat.search(name):
            continue
        print(f"== {name}: {len(ins)} instructions")
        idx = {a: k for k, (a, _, _) in enumerate(ins)}
        for k, (addr, text, hi) in enumerate(ins):
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt > addr or tgt not in idx:
                continue
            body = ins[idx[tgt]: k + 1]
            stalls = sum(ctrl(h)["stall"] for _, _, h in body)
            waits = sum(1 for _, _, h in body if ctrl(h)["wait"])
            hist = collections.Counter(opcode(t) for _, t, _ in body)
            top = ", ".join(f"{o} {c}" for o, c in hist.most_common(14))
            print(f"  loop 0x{tgt:04x}..0x{addr:04x}: {len(body)} instr, stall sum {stalls}, {waits} with scoreboard waits | {top}")
            if dump:
                for a, t, h in body:
                    c = ctrl(h)
                    print(f"      {a:04x}  s{c['stall']:<2d} {'Y' if c['yield'] else ' '} w{c['wbar']} r{c['rbar']} m{c['wait']:02x}  {t}")


if __name__ == "__main__":
    main()
