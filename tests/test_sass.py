"""The built library is Blackwell-native: one sm_100a image, and the dense-layer kernels issue tcgen05 / TMEM / TMA
instructions (no mma.sync fall-back).  CPU test: reads the SASS of the in-tree libhgin.so with cuobjdump, launches nothing.
The mnemonics are the ones /opt/skills/guides/B200_PROFILING.md lists as the proof of tcgen05 (UTCHMMA), tcgen05.ld
(LDTM), TMA tensor copies (UTMALDG / UTMASTG) and tcgen05.commit (UTCBAR)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gnn_link_prediction_b200", "libhgin.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(SO),
                                reason="needs cuobjdump and the built libhgin.so (__graft_entry__.build())")


@pytest.fixture(scope="module")
def sass():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            while cur in kernels:
                cur += "'"
            kernels[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P(?:\d+|T)\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            kernels[cur].append(m.group(1))
    return set(re.findall(r"arch = (sm_\w+)", out)), kernels


def count(ops, prefix):
    return sum(op.startswith(prefix) for op in ops)


def test_library_holds_one_sm_100a_image(sass):
    archs, kernels = sass
    assert archs == {"sm_100a"}
    assert len(kernels) > 100


@pytest.mark.parametrize("name", ["gemm_nt_kernel", "gemm_tn_kernel", "gemm_nt_bf16_kernel", "gemm_tn_bf16_kernel"])
def test_dense_layer_kernels_run_on_tcgen05_with_tma_operands(sass, name):
    _, kernels = sass
    mine = {k: ops for k, ops in kernels.items() if name in k}
    assert mine, name
    for k, ops in mine.items():
        assert count(ops, "UTCHMMA") >= 4, (k, "tcgen05.mma")
        assert count(ops, "LDTM") >= 1, (k, "tcgen05.ld of the TMEM accumulator")
        assert count(ops, "UTMALDG") >= 1, (k, "TMA tensor loads of the operands")
        assert count(ops, "UTCBAR") >= 1, (k, "tcgen05.commit")
        assert count(ops, "HMMA") == 0, (k, "no mma.sync path")


def test_no_kernel_uses_legacy_tensor_core_instructions(sass):
    _, kernels = sass
    legacy = [k for k, ops in kernels.items() if any(op.startswith(("HMMA", "IMMA", "HGMMA")) for op in ops)]
    assert legacy == []


def test_aggregation_kernels_are_simt_vector_loads_without_local_memory(sass):
    """K1/K4 are HBM-bound gathers: 128-bit global loads, no spills (LDL/STL) in the default (global-gather) variants."""
    _, kernels = sass
    agg = {k: ops for k, ops in kernels.items() if "gin_combine_kernel" in k}
    assert len(agg) >= 20
    assert sum(count(ops, "LDG.E.128") + count(ops, "LDG.E.EL.128") + count(ops, "LDG.E.EF.128")
               + count(ops, "LDG.E.128.CONSTANT") for ops in agg.values()) > 0
    spilled = [k for k, ops in agg.items() if count(ops, "STL") > 0]
    assert len(spilled) <= len(agg) // 10, spilled[:5]
