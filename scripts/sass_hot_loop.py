"""Static view of the pairwise fast loop of the built library (no GPU needed):

    python scripts/sass_hot_loop.py > profiles/rNN_fast_loop_sass.txt

Disassembles ssa_pairwise_kernel<3, false, true, true> with cuobjdump, finds the inner loop (the backward branch
that closes the CT_BLOCKS_PER_CHECK loop over Philox blocks), walks its main path -- the path of a warp in which
no lane crosses a grid point: forward branches over the grid-crossing block are followed -- and prints the
instruction count per SSA step, the opcode / pipe mix and the listing."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "circuitree_b200" / "libcircuitree_b200.so"
KERNEL = "_ZN2ct19ssa_pairwise_kernelILi3ELb0ELb1ELb1EEEvNS_10LaunchArgsE"

# pipe of an opcode (B200_PROFILING.md / ncu pipe names): FMA-pipe ops issue at full rate, ALU-pipe ops at half rate
FMA = {"FFMA", "FMUL", "FADD", "IMAD", "HFMA2"}
ALU = {"LOP3", "FSEL", "FSETP", "FMNMX", "ISETP", "IADD3", "VIADD", "LEA", "PRMT", "PLOP3", "SEL", "SHF", "MOV", "VIMNMX", "FSET", "I2FP", "F2I"}
XU = {"MUFU", "FRND"}


def main():
    lib = Path(sys.argv[1]) if len(sys.argv) > 1 else LIB
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", KERNEL, str(lib)], capture_output=True, text=True).stdout
    ins = []
    for line in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    if not ins:
        sys.exit("kernel not found in " + str(lib))
    index = {a: i for i, (a, _) in enumerate(ins)}
    # inner loop: the backward conditional branch with the longest span
    best = None
    for i, (a, s) in enumerate(ins):
        m = re.match(r"@!?U?P\d BRA(?:\.U)? (0x[0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16)
            if t < a and (best is None or a - t > best[1] - best[0]) and "FFMA.SAT" in " ".join(x for _, x in ins[index[t]:i]):
                best = (t, a)
    start, end = index[best[0]], index[best[1]]
    path, i = [], start
    while i <= end:
        a, s = ins[i]
        path.append((a, s))
        m = re.match(r"@!?U?P\d BRA(?:\.U)? (0x[0-9a-f]+)", s)
        if m and i != end:
            t = int(m.group(1), 16)
            if t > a and index.get(t, 1 << 30) <= end and t - a > 0x300:      # over the grid-crossing block
                i = index[t]
                continue
        i += 1
    ops = collections.Counter()
    pipes = collections.Counter()
    for _, s in path:
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        base = op.split(".")[0]
        ops[base] += 1
        pipes["fma" if base in FMA else "alu" if base in ALU else "xu" if base in XU else "other"] += 1
    n = len(path)
    print(f"# {lib.name}: ssa_pairwise_kernel<3, 0, 1, 1>, inner loop {best[0]:#x}..{best[1]:#x} ({end - start + 1} instructions in the text)")
    print(f"# main path (no lane crosses a grid point): {n} instructions per Philox block = {n / 2:.1f} per SSA step")
    print(f"# pipes: " + ", ".join(f"{k} {v}" for k, v in pipes.most_common()))
    print(f"# opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common()))
    crossing = [s for a, s in ins[start:end + 1] if (a, s) not in set(path)]
    print(f"# off the main path (grid crossing, recording, finishing): {len(crossing)} instructions")
    print()
    for a, s in path:
        print(f"{a:05x}  {s}")


if __name__ == "__main__":
    main()
