"""Instruction histogram of selected kernels from `cuobjdump -sass` (profiles/r2_sass_operator_kernels.txt).
    python tools/sass_summary.py topomax_b200/libtopomax_b200.so elast_apply_kernelIdLb0ELi3ELi2ELb1 ..."""
import collections
import re
import subprocess
import sys


def main(lib, names):
    lines = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
    for name in names:
        starts = [i for i, l in enumerate(lines) if "Function :" in l and name in l]
        if not starts:
            print(f"== {name}: not found")
            continue
        s = starts[0]
        e = next((i for i in range(s + 1, len(lines)) if "Function :" in lines[i]), len(lines))
        ops = collections.Counter()
        for l in lines[s:e]:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                ops[m.group(2)] += 1
        groups = collections.Counter()
        for k, v in ops.items():
            groups[k.split(".")[0]] += v
        print(f"== {name}: {sum(ops.values())} SASS instructions")
        print("   by opcode: " + ", ".join(f"{k} {v}" for k, v in groups.most_common(24)))
        mem = {k: v for k, v in ops.items() if k.startswith(
            ("LDG", "STG", "LDS", "STS", "LDL", "STL", "SHFL", "CCTL", "LDC", "ATOM", "RED", "BAR", "UTMA", "UBLK",
             "LDGSTS", "SYNCS"))}
        print("   memory / exchange instructions: " + ", ".join(f"{k} {v}" for k, v in sorted(mem.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
