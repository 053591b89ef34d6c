"""Static SASS instruction mix of one kernel in an object file (build box, no GPU).

usage: python scripts/sass_mix.py <file.o|.so> <kernel-name-substring> [--dump]

Splits the kernel's code at RET instructions into the caller body and its out-of-line callees (ptxas places them after
the body), classifies every instruction by issue pipe and prints the counts that matter for the multiplier-bound
ladders: IMAD.WIDE (the work), everything else that issues on the same FMA pipe (IMAD.MOV, IMAD, IMAD.X, IMAD.IADD,
IMAD.HI ...), ALU-pipe instructions, and the register moves per call site."""
import collections
import re
import subprocess
import sys


def pipe(op):
    if op.startswith("IMAD.WIDE"):
        return "wide"
    if op.startswith(("IMAD", "FFMA", "FMUL", "FADD", "HFMA", "DFMA")):
        return "fma_other"
    if op.startswith(("LD", "ST", "ATOM", "RED", "LDS", "STS", "LDG", "STG", "LDL", "STL", "LDC")):
        return "lsu"
    if op.startswith(("BRA", "CALL", "RET", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "NOP", "BAR", "JMP", "BRX")):
        return "control"
    return "alu"


def kernel_sass(path, name):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out, on = [], False
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            on = name in m.group(1)
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins = m.group(2).strip()
                ins = re.sub(r"^@!?U?P\d\s+", "", ins)
                out.append((int(m.group(1), 16), ins))
    return out


def main():
    path, name = sys.argv[1], sys.argv[2]
    ins = kernel_sass(path, name)
    if not ins:
        sys.exit("kernel not found")
    # call targets -> function starts
    targets = sorted({int(m.group(1), 16) for _, i in ins for m in [re.match(r"CALL\.\S+ (0x[0-9a-f]+)", i)] if m})
    bounds = [0] + targets + [ins[-1][0] + 16]
    print(f"{name}: {len(ins)} instructions, {len(targets)} out-of-line callees")
    for k in range(len(bounds) - 1):
        seg = [(a, i) for a, i in ins if bounds[k] <= a < bounds[k + 1]]
        ops = [i.split()[0] for _, i in seg]
        # trailing self-branch / NOP padding after the last RET / EXIT does not execute
        cnt = collections.Counter(pipe(o) for o in ops if o != "NOP")
        top = collections.Counter(ops).most_common(14)
        label = "caller" if k == 0 else f"callee@{bounds[k]:#x}"
        print(f"== {label}: {len(seg)} instr  wide={cnt['wide']} fma_other={cnt['fma_other']} alu={cnt['alu']} lsu={cnt['lsu']} control={cnt['control']}")
        print("   ", top)
        if k == 0:
            calls = collections.Counter(int(m.group(1), 16) for _, i in seg for m in [re.match(r"CALL\.\S+ (0x[0-9a-f]+)", i)] if m)
            nmov = sum(1 for o in ops if o.startswith("IMAD.MOV")) + sum(1 for o in ops if o == "MOV")
            print(f"    call sites: {dict((hex(a), n) for a, n in calls.items())}; register moves in caller: {nmov} "
                  f"({nmov / max(1, sum(calls.values())):.1f} per call)")
        if "--dump" in sys.argv and k > 0:
            for a, i in seg:
                print(f"      {a:05x}  {i}")


if __name__ == "__main__":
    main()
