"""Per-kernel counts of the SASS instructions that show the library is Blackwell-native (CPU only):
    python profiles/sass_summary.py > profiles/sass_r2.txt
UBLKCP = TMA bulk copies (cp.async.bulk), SYNCS = mbarrier waits, ATOMG.E.CAS.128 = 128-bit claims,
REDG / ATOMG = fire-and-forget reductions / atomics with a return, ATOMS = shared-memory atomics."""
import os
import re
import subprocess
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vdjer_b200", "libvdjgraph.so")
PATTERNS = OrderedDict([("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("ATOMG.CAS.128", r"ATOMG\S*CAS\S*128"), ("ATOMG", r"\bATOMG"),
                        ("REDG", r"\bREDG"), ("ATOMS", r"\bATOMS"), ("LDG.256", r"LDG\S*\.256"), ("STG.256", r"STG\S*\.256"),
                        ("SHFL", r"\bSHFL"), ("VOTE", r"\bVOTE"), ("STL/LDL", r"\b(STL|LDL)\b")])


def main():
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    print(f"# cuobjdump -sass vdjer_b200/libvdjgraph.so  (architectures in the fat binary: {', '.join(arch)})")
    print("kernel".ljust(34) + " ".join(k.rjust(13) for k in PATTERNS) + "   instructions")
    cur, body = None, []

    def flush():
        if cur is None:
            return
        text = "\n".join(body)
        n = len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", text, re.M))
        print(cur[:33].ljust(34) + " ".join(str(len(re.findall(p, text))).rjust(13) for p in PATTERNS.values()) + f"   {n}")

    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            name = m.group(1)
            d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            k = re.search(r"(?:vdjg::)?(k_\w+)(<[^>]*>)?", d)
            cur = (k.group(1) + (k.group(2) or "")) if k else d[:33]
            body = []
        elif cur is not None:
            body.append(line)
    flush()


if __name__ == "__main__":
    main()
