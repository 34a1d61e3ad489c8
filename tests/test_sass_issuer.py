"""CPU: the decoder's MMA issue loop stays lean in the BUILT library (SASS of libneuralbody_b200.so, via cuobjdump).

Round 2 found the issuer's own instruction stream -- not the tensor pipe -- bounding the short layers: under `if (lane == 0)`
ptxas wraps every uniform-register operand of a UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop (~14 SASS
instructions per MMA).  The warp-uniform `elect_one()` form (nb_tc_ptx.cuh) emits back-to-back UTCHMMA.  This test pins that
property of the shipped binary, so a refactor that silently re-introduces the waterfall shows up without a GPU."""
import re
import shutil
import subprocess

import pytest

from neuralbody_b200 import _build


def _functions():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not shutil.os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    lib = _build.build()
    txt = subprocess.run([exe, "-sass", lib], capture_output=True, text=True, check=True).stdout
    out = {}
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name, _, body = part.partition("\n")
        out[name.strip()] = body
    return out


def test_decoder_issue_loop_is_warp_uniform():
    fns = {n: b for n, b in _functions().items() if "render_tc_list_kernel" in n}
    assert len(fns) == 4, sorted(fns)            # <1 | 3 passes> x <fp32 | fp16 volume>
    for name, body in fns.items():
        mma = len(re.findall(r"\bUTCHMMA\b", body))
        waterfall = len(re.findall(r"BRA\.U\.ANY", body))
        commits = len(re.findall(r"\bUTCBAR\b", body))
        assert ".2CTA" in body                   # CTA pairs: tcgen05 cta_group::2
        # every K-step of every layer is its own (unrolled) MMA site: 3-pass 18 + 1 + 2 x 49 + 22 = 139, 1-pass 6 + 1 + 2 x 17 + 22 = 63
        assert mma >= (139 if "ILi3E" in name else 63), (name, mma)
        assert commits >= 10, (name, commits)
        # the single-lane form had one waterfall loop per MMA (and per commit); the warp-uniform form keeps a handful in cold
        # divergence-fallback paths only
        assert waterfall * 4 < mma, "%s: %d BRA.U.ANY loops for %d UTCHMMA -- the issuer fell back to the single-lane form" % (name, waterfall, mma)
