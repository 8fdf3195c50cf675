"""Build sanity of the shipped library, checked without a GPU: every tcgen05 kernel of the extension still contains its
tensor-core and TMA instructions in SASS.  (A preprocessor slip once compiled the MMA issuer loop out of the persistent
kernel: everything built, nothing computed, and the first launch hung.)"""
import os
import re
import shutil
import subprocess

import pytest

from point_unet_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass_by_function():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    text = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    out, name = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name is not None:
            out[name].append(line)
    return {k: "\n".join(v) for k, v in out.items()}


def _count(body, mnemonic):
    return len(re.findall(r"\b" + mnemonic, body))


def test_tensor_core_kernels_issue_mma(sass_by_function):
    persist = {k: v for k, v in sass_by_function.items() if "tc_persist_kernel" in k}
    wgrad = {k: v for k, v in sass_by_function.items() if "tc_wgrad" in k}
    assert len(persist) >= 6 and len(wgrad) >= 2, (sorted(persist), sorted(wgrad))
    for name, body in {**persist, **wgrad}.items():
        assert _count(body, "UTCHMMA") >= 3, f"{name}: no tcgen05.mma left in SASS"
        assert _count(body, "UTCBAR") >= 1, f"{name}: no tcgen05.commit"
        assert _count(body, "SYNCS") >= 4, f"{name}: no mbarrier traffic"
    # the TMA-fed kernels keep their bulk copies; the one-pass att backward keeps BOTH of its MMA sites
    assert any(_count(b, "UTMALDG") or _count(b, "UBLKCP") for b in persist.values())
    fused = [b for k, b in persist.items() if "Li64ELi3ELb0" in k]
    plain = [b for k, b in persist.items() if "Li64ELi2ELb0" in k]
    assert fused and plain
    assert _count(fused[0], "UTCHMMA") > _count(plain[0], "UTCHMMA")


def test_sm100a_only(sass_by_function):
    text = subprocess.run([CUOBJDUMP, "-lelf", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", text))
    assert archs == {"100a"}, archs
