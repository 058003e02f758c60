"""Structural checks on the built libpgpd.so (no GPU needed: cuobjdump disassembles the sm_100a cubin).

* the tensor-core path really is tcgen05 / TMEM / bulk + tensor-map copies (UTCHMMA, LDTM, UBLKCP, UTMALDG in the SASS);
* every tcgen05 kernel issues its MMAs from warp-uniform code: issued under `if (lane == 0)` the compiler wraps EVERY tcgen05.mma in
  an ELECT / R2UR.BROADCAST x5 / branch waterfall (round 2 measured ~130 cycles of issue per MMA that way; `k_kf_tc` had 187 R2UR for
  its 12 UTCHMMA).  With uniform operands (tc_ptx.cuh: warp_uniform, elect_one) the R2UR count stays below the UTCHMMA count."""
import os
import re
import shutil
import subprocess

import pytest

from pointnetgpd_b200 import _abi as A

KERNELS = ["k_l3_fwd_tc3", "k_tower_fused_eval", "k_kf_tc", "k_ka_tc", "k_kb_tc", "k_gemm_tc"]


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(A.LIB_PATH):
        pytest.skip("libpgpd.so not built")
    out = subprocess.run([exe, "-sass", A.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    assert funcs, "no functions in the disassembly"
    return funcs


def _body(funcs, kernel):
    hits = [v for k, v in funcs.items() if kernel in k]
    assert hits, "kernel %s not found in libpgpd.so" % kernel
    return "\n".join(hits[0])


def test_tensor_core_path_is_tcgen05(sass):
    l3 = _body(sass, "k_l3_fwd_tc3")
    assert "UTCHMMA.2CTA" in l3            # tcgen05.mma.cta_group::2
    assert "LDTM" in l3                    # tcgen05.ld: accumulators live in TMEM
    assert "UTMALDG" in l3                 # tensor-map copies of the weight stream
    assert "UBLKCP" in _body(sass, "k_ka_tc")   # bulk copies of the raw operand rows
    assert "HMMA" not in l3.replace("UTCHMMA", "")   # no mma.sync / wmma fallback in the dominant kernel


@pytest.mark.parametrize("kernel", KERNELS)
def test_mma_issue_is_warp_uniform(sass, kernel):
    body = _body(sass, kernel)
    n_mma, n_r2ur = body.count("UTCHMMA"), body.count("R2UR")
    assert n_mma >= 12, (kernel, n_mma)
    assert n_r2ur < n_mma + 16, "%s: %d R2UR for %d UTCHMMA -- the MMAs are being issued from per-thread registers again" % (kernel, n_r2ur, n_mma)
