"""SASS evidence for a kernel of libffx.so: `cuobjdump -sass` of the shipped sm_100a cubin, cut to
one kernel, as an opcode histogram plus the lines that prove the bulk-copy / mbarrier path.

    python tools/sass_summary.py "ffx_score_tma_kernelILi2ELi12ELb1ELi32E" > profiles/r2_score_tma_sass.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fast-forward-indexes_b200", "fast_forward", "lib", "libffx.so")


def main():
    want = sys.argv[1]
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    blocks = re.split(r"\n\s*Function : ", text)
    picked = [b for b in blocks[1:] if want in b.split("\n", 1)[0]]
    if not picked:
        sys.exit(f"no kernel matching {want}")
    for b in picked:
        name, body = b.split("\n", 1)
        demangled = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip()
        ops = collections.Counter()
        keep = []
        for line in body.splitlines():
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if not m:
                continue
            op = m.group(1)
            ops[op] += 1
            if op.startswith(("UBLKCP", "SYNCS", "UTMA", "FENCE", "ATOMS", "REDUX", "MATCH")):
                keep.append(line.split("/*", 2)[1][-5:-1] + "  " + line.split("*/", 1)[1].split("/*")[0].strip())
        print(f"# {demangled}\n# mangled: {name.strip()}\n# source: cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a)")
        print(f"# instructions: {sum(ops.values())}")
        fam = collections.Counter()
        for op, n in ops.items():
            fam[op.split(".")[0]] += n
        print("\n## opcode families")
        for op, n in fam.most_common():
            print(f"{n:6d}  {op}")
        print("\n## full opcodes")
        for op, n in sorted(ops.items(), key=lambda kv: (-kv[1], kv[0])):
            print(f"{n:6d}  {op}")
        print("\n## bulk-copy engine / mbarrier / fence / warp-match lines")
        for ln in keep:
            print(ln)
        tensor = [op for op in ops if op.startswith(("HMMA", "UTC", "TCGEN", "UMMA", "IMMA", "DMMA", "QMMA"))]
        print(f"\n## tensor-core / TMEM opcodes: {tensor or 'none (ragged gather-GEMV: HBM-bound, as north_star prescribes)'}")
        print()


if __name__ == "__main__":
    main()
