"""Fingerprint of the dense fused-pair kernels' SASS (instruction count + hash of the opcode/operand stream).

The headline kernel pair_rows_persistent<float,2,PAIR_ROW> moves by +-2 % with ptxas' schedule of the same algorithm
(DESIGN.md section 3), so a change to ofd_pair.cu that is not meant to touch the dense path should leave this fingerprint
unchanged.  Usage: python tools/pair_sass_fingerprint.py [--check profiles/r1/pair_sass_fingerprint.txt]
Runs on the CPU (needs the built object opticalflowfromdepth_b200/build/ofd_pair.cu.o and cuobjdump)."""
import hashlib
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OBJ = ROOT / "opticalflowfromdepth_b200" / "build" / "ofd_pair.cu.o"


def fingerprints():
    sass = subprocess.run(["cuobjdump", "-sass", str(OBJ)], stdout=subprocess.PIPE, text=True, check=True).stdout
    out, cur = {}, None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1) if "pair_rows_persistent" in m.group(1) else None
            if cur:
                out[cur] = []
            continue
        if cur and re.search(r"/\*[0-9a-f]{4}\*/", ln):
            ins = re.sub(r"/\*.*?\*/", "", ln).strip()
            if ins:
                out[cur].append(re.sub(r"\s+", " ", ins))
    res = {}
    for name, ins in out.items():
        m = re.search(r"persistentI(.)Li(\d)ELi(\d)E", name)
        key = f"pair_rows_persistent<{'float' if m.group(1) == 'f' else 'double'},{m.group(2)},{('PAIR_ROW', 'PAIR_GROUPED', 'PAIR_RAGGED', 'PAIR_RAGGED_ANY')[int(m.group(3))]}>"
        res[key] = (len(ins), hashlib.sha256("\n".join(ins).encode()).hexdigest()[:16])
    return res


def main():
    fp = fingerprints()
    lines = [f"{k:55s} {n:5d} instructions  {h}" for k, (n, h) in sorted(fp.items())]
    if len(sys.argv) > 2 and sys.argv[1] == "--check":
        want = {ln.split()[0]: ln.split()[-1] for ln in Path(sys.argv[2]).read_text().splitlines() if ln.startswith("pair_rows")}
        bad = [k for k, (_, h) in fp.items() if "RAGGED" not in k and want.get(k) != h]
        print("\n".join(lines))
        print("dense kernels unchanged" if not bad else f"CHANGED: {bad} - re-measure the headline before committing")
        return 1 if bad else 0
    print("\n".join(lines))
    return 0


if __name__ == "__main__":
    sys.exit(main())
