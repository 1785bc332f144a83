"""CPU: static check of the built library's SASS.  A tcgen05.mma whose operands are not provably warp-uniform is issued
through ELECT + R2UR.BROADCAST moves, which profiles/r01_gemm_issue.md measured at ~3x the issue time of an MMA fed from
uniform registers (and the narrow-tile kernels are issue-bound).  Things that silently break uniformity for a whole
kernel: a named barrier in another warp role, a 64-bit division in the item decode, operands derived from LDS results or
from threadIdx-based warp indices.  This test keeps every tensor-core kernel (default and opt-in variants) clean."""
import os
import re
import shutil
import subprocess
from collections import Counter

import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vitta_b200", "libvitta_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.mark.skipif(not os.path.exists(CUOBJDUMP) or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_tensor_core_kernels_issue_from_uniform_registers():
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur is not None:
            t = m.group(1).split()
            cur[t[1] if t[0].startswith("@") else t[0]] += 1
    tc = {k: c for k, c in kernels.items() if c.get("UTCHMMA", 0) + c.get("UTCHMMA.2CTA", 0) > 0}
    assert len(tc) >= 20, "expected the GEMM / wgrad / attention kernels and their opt-in variants"
    for name, c in tc.items():
        mma = c.get("UTCHMMA", 0) + c.get("UTCHMMA.2CTA", 0)
        # a handful of moves (predicates of ragged k-steps in the attention backward) are tolerated; the pathological
        # state is ~3 per MMA
        assert c.get("R2UR.BROADCAST", 0) <= 8, (name, mma, c.get("R2UR.BROADCAST", 0))


@pytest.mark.skipif(not os.path.exists(CUOBJDUMP) or not os.path.exists(LIB), reason="needs cuobjdump and the built library")
def test_streaming_kernels_fit_their_register_budget_without_spills():
    """The HBM-bound kernels are sized for a fixed number of resident CTAs per SM (their grids and the loads they keep in
    flight assume it): the default-path instantiations must not spill to local memory, and the sub-warp LayerNorm kernels
    must stay inside the 128 registers two resident CTAs allow."""
    out = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True, timeout=600).stdout
    usage, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
            cur = None
    assert len(usage) > 100
    want = ["bn_relu_pool_bwd_kernel", "bn_relu_pool_fwd_kernelIj", "tam_bwd_kernel", "tam_fwd_kernel", "ln_fwd_narrow_kernel",
            "ln_bwd_narrow_kernel", "stats_cl_kernel", "bn_act_fwd_kernel", "bn_act_bwd_amax_kernel"]
    # (ln_bwd_split_kernel<8> is known to keep 32 bytes of stack at C > 1536 and is not listed)
    for key in want:
        hits = {k: v for k, v in usage.items() if key in k}
        assert hits, key
        for name, (regs, stack) in hits.items():
            assert stack == 0, (name, regs, stack)
            if "narrow" in name:
                assert regs <= 128, (name, regs)
