"""Kernels launched with the programmatic-dependent-launch attribute start before their predecessor has finished and hold at
griddepcontrol.wait (ACQBULK).  Compiler and assembler both move read-only loads (__ldg / const __restrict__ = ld.global.nc) above that
wait if nothing stops them; such a load reads the previous frame's data (found with the queue length of sph3_force_heavy_kernel).
CWA_PDL_ENTER makes the kernel body control-dependent on the wait; this test checks the result where it counts, in the SASS of every
built object: no memory instruction before the ACQBULK of any kernel."""
import glob
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_no_kernel_touches_memory_before_its_dependency_wait():
    from coupledwateranimation_b200 import build as B
    import check_pdl_sass
    B.build()
    objs = sorted(glob.glob(os.path.join(ROOT, "coupledwateranimation_b200", "build", "*.o")))
    assert objs, "no objects built"
    waiting = 0
    for o in objs:
        seen, bad = check_pdl_sass.check(o)
        waiting += seen
        assert not bad, f"{os.path.basename(o)}: memory access before griddepcontrol.wait in {[b[0][:60] for b in bad]}: {bad[0][1][:4]}"
    assert waiting >= 20, f"only {waiting} kernels with a dependency wait found: the check is not looking at the right files"


def test_every_kernel_launched_with_the_attribute_waits_first():
    """cwa_launch may attach the attribute to any kernel it is given: each of them must begin with CWA_PDL_ENTER (a kernel without the wait
    would run on its predecessor's half-written output)."""
    import re
    csrc = os.path.join(ROOT, "coupledwateranimation_b200", "csrc")
    text = {f: open(os.path.join(csrc, f)).read() for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))}
    launched = set()
    for t in text.values():
        for m in re.finditer(r"cwa_launch\(ctx,\s*[^,]+,\s*([A-Za-z0-9_]+)\s*(?:<[^>]*>)?\s*,", t):
            launched.add(m.group(1))
    launched.discard("kern")                              # the helper's own parameter name
    assert len(launched) >= 12, launched
    for k in sorted(launched):
        bodies = [m for t in text.values() for m in re.finditer(r"\b" + k + r"\([^;{]*\)\s*\{\s*\n\s*(\S+)", t)]
        assert bodies, f"definition of {k} not found"
        assert all(b.group(1).startswith("CWA_PDL_ENTER()") for b in bodies), f"{k} does not begin with CWA_PDL_ENTER()"
