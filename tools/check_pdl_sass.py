"""Every kernel that waits on its predecessor (griddepcontrol.wait = ACQBULK in the SASS) must not touch global memory before the wait:
the compiler may hoist read-only (ld.global.nc / const __restrict__) loads above an inline-asm barrier, and a load that runs before the
predecessor grid has finished reads what the previous frame left there.  Lists, per object file, the kernels with an ACQBULK and the
memory instructions that precede it (none allowed).  usage: python tools/check_pdl_sass.py [objects...]"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEM = re.compile(r"\b(LDG|LD|STG|ST|ATOM|ATOMG|RED|LDGSTS|UBLKCP|UTMALDG|CCTL)\b")


def check(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    bad, seen = [], 0
    name, before, waited = None, [], False
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name, before, waited = m.group(1), [], False
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m or name is None or waited:
            continue
        ins = m.group(2)
        if "ACQBULK" in ins:
            waited = True
            seen += 1
            if before:
                bad.append((name, before))
        elif MEM.search(ins.split()[0] if not ins.startswith("@") else ins.split()[1]):
            before.append(ins.strip())
    return seen, bad


if __name__ == "__main__":
    objs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "coupledwateranimation_b200", "build", "*.o")))
    rc = 0
    for o in objs:
        seen, bad = check(o)
        print(f"{os.path.basename(o)}: {seen} kernels wait on their predecessor, {len(bad)} touch memory before the wait")
        for name, ins in bad:
            rc = 1
            print("   ", name[:90])
            for i in ins[:6]:
                print("        ", i)
    sys.exit(rc)
