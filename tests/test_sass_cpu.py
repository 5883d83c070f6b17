"""CPU: the built library is Blackwell-native.  cuobjdump (no GPU needed) must show the tcgen05 / TMEM / TMA mnemonics in
the implicit-GEMM kernels -- UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG / UTMAREDG
(cp.async.bulk.tensor load / store / reduce) -- and no warp-level HMMA / IMMA (mma.sync, wmma) anywhere in the library;
the folded-GroupNorm instantiations must carry the packed f32x2 FMAs and MUFU.TANH of their transform warps."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "masked-diffusion-model_b200", "mdm_b200", "libmdm_sm100.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    if not os.path.exists(LIB):
        pytest.skip("library not built")
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    per, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = []
        elif cur is not None:
            per[cur].append(line)
    assert per, "no SASS found (was the library built for sm_100a?)"
    return {k: "\n".join(v) for k, v in per.items()}


def test_gemm_kernels_use_tcgen05_tmem_and_tma(sass):
    ig = {k: v for k, v in sass.items() if "igemm_kernel" in k}
    assert len(ig) >= 20, len(ig)
    for name, body in ig.items():
        assert "UTCHMMA" in body, name                      # tcgen05.mma
        assert "LDTM" in body, name                         # tcgen05.ld (accumulators out of TMEM)
        assert "UTMALDG" in body, name                      # TMA tensor loads
        assert "UTMASTG" in body or "UTMAREDG" in body, name   # TMA stores / reduce-adds


def test_no_legacy_warp_level_mma_anywhere(sass):
    for name, body in sass.items():
        assert not re.search(r"\bHMMA\b|\bIMMA\b", body), name


def test_folded_groupnorm_instantiations_carry_the_transform_code(sass):
    # template arguments <mode, halo, MT, stats, dyn, norm>: the last one is Lb1 for the folded variants
    folded = {k: v for k, v in sass.items() if "igemm_kernel" in k and k.count("Lb1E") and re.search(r"Lb[01]ELb[01]ELb1EE", k)}
    assert len(folded) == 4, sorted(folded)
    for name, body in folded.items():
        assert body.count("MUFU.TANH") >= 88, name          # 11 pieces x 8 channels per thread and chunk
        assert "FFMA2" in body, name                        # fma.rn.f32x2
        assert "BAR.SYNC.DEFER_BLOCKING 0x2, 0x100" in body, name   # the 256 transform threads' own named barrier
