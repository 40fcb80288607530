"""Static checks of the built library's SASS (no GPU needed): the properties the measured speed depends on and that a small
source edit can silently flip, because ptxas chooses the instruction forms.

 * weights reach the FMA pipe through the uniform datapath (LDCU, immediate or uniform-register offset).  Per-thread
   register-indexed constant loads (LDC c[0x0][Rn+imm]) go through the indexed constant cache, whose ~0.5 requests/clk/SM
   bound the decoder 3x3 kernel at 9.1 ms instead of 7.5 ms and the GTCN kernel at 8.7 ms instead of 4.5 ms (B200, round 1).
 * the convolutions are packed FFMA2 (fma.rn.f32x2), not scalar FFMA.
 * no kernel on the default path spills more than a couple of registers (<= 16 bytes of stack).
"""
import os
import re
import shutil
import subprocess

import pytest

import gtcrn_micro_b200.build as B

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
# kernels launched with the default tuning (gtcrn_capi.cu): offline + streaming
HOT = ["k_frontILb0ELi0ELb1E", "k_frontILb1ELi0ELb1E", "k_frontILb0ELi0ELb0E", "k_frontILb1ELi0ELb0E",
       "k_backILb0ELi0ELb1E", "k_backILb1ELi0ELb1E", "k_backILb0ELi0ELb0E", "k_backILb1ELi0ELb0E",
       "k_gtconvILb0ELi2ELi4ELb0", "k_gtconvILb1ELi4ELi7ELb0", "k_gtcnILi2ELb0", "ks_gtconvILb0", "ks_gtconvILb1", "ks_gtconv_enc3", "ks_gtcnE"]
# tensor-core kernels of the default path (gtcrn_tc.cuh): decoder GTConv with 4 warpgroups per CTA, GTCN with one
HOT_TC = ["k_gtconv_dec_tcILi4ELb0", "k_gtconv_dec_tcILi4ELb1", "k_gtcn_tcILi1ELb0"]
# en1 inside k_front on the tensor cores (default): both operands come from shared memory (descriptors), the accumulator from TMEM
HOT_TC_SS = ["k_frontILb0ELi0ELb1E", "k_frontILb1ELi0ELb1E", "k_backILb0ELi0ELb1E", "k_backILb1ELi0ELb1E"]
# kernels of the original-GTCRN-style variant (BASELINE configs[3]): same weight-feed property, spills tolerated up to a small frame
HOT_FULL = ["k_frontILb0ELi1ELb0E", "k_frontILb1ELi1ELb0E", "k_backILb0ELi1ELb0E", "k_backILb1ELi1ELb0E", "kf_gtconvILb0ELi2ELi4E", "kf_gtconvILb1ELi2ELi4E",
            "kf_dp_intra", "kf_dp_inter"]


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    lib = B.build()
    out = subprocess.run([CUOBJDUMP, "-sass", lib], capture_output=True, text=True, check=True).stdout
    per, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1); per[name] = []
        elif name and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
            per[name].append(line)
    return per


def _find(per, key):
    hits = [n for n in per if key in n]
    assert len(hits) == 1, (key, hits)
    return per[hits[0]]


@pytest.mark.parametrize("key", HOT + HOT_FULL)
def test_weights_use_the_uniform_datapath(sass, key):
    body = _find(sass, key)
    r_indexed = sum(1 for l in body if re.search(r"c\[0x0\]\[R\d", l))
    limit = 48 if key == "ks_gtconv_enc3" else 16                 # three GTConvBlocks in one kernel (the TRALite gate's 9 per block)
    assert r_indexed <= limit, f"{key}: {r_indexed} per-thread-indexed constant loads (LDC c[0x0][R..]) -- see tools/sass_check.sh"


@pytest.mark.parametrize("key", HOT_TC)
def test_tensor_core_kernels_use_tcgen05_with_operands_in_tensor_memory(sass, key):
    """The TC kernels must contain the tcgen05 MMA (SASS UTCHMMA), write their A operand with tcgen05.st (STTM) and read the
    accumulator with tcgen05.ld (LDTM); weights still come through the uniform datapath; no spills."""
    body = _find(sass, key)
    n = {op: sum(1 for l in body if re.search(r"\s" + op + r"[.\s]", l)) for op in ("UTCHMMA", "STTM", "LDTM", "UTCBAR")}
    assert n["UTCHMMA"] >= 6 and n["STTM"] >= 2 and n["LDTM"] >= 2 and n["UTCBAR"] >= 1, (key, n)
    # A comes from tensor memory: the MMA's second operand is tmem[..], not a shared-memory descriptor
    mma = [l for l in body if "UTCHMMA" in l]
    assert all(len(re.findall(r"tmem\[", l)) >= 3 for l in mma), mma[:2]
    r_indexed = sum(1 for l in body if re.search(r"c\[0x0\]\[R\d", l))
    assert r_indexed <= 16, (key, r_indexed)
    assert not any(re.search(r"\s(STL|LDL)[.\s]", l) for l in body), f"{key} spills"


@pytest.mark.parametrize("key", HOT_TC_SS)
def test_ends_run_en1_and_de3_on_tcgen05_from_shared_memory(sass, key):
    """k_front: 5 taps x (hi' [W_hi | W_lo], lo W_hi) = 10 MMAs in the (rolled) tile loop; k_back: 3 tap pairs x 3 = 9.  A and B through
    shared-memory descriptors, the accumulator read back with tcgen05.ld; the fp32 en1 / de3 FFMA loops are gone (en0's 240 and
    de4's FFMA2 remain); tables and tiles arrive by TMA bulk copies (UBLKCP)."""
    body = _find(sass, key)
    n = {op: sum(1 for l in body if re.search(r"\s" + op + r"[.\s]", l)) for op in ("UTCHMMA", "LDTM", "UTCBAR", "FFMA2", "UBLKCP")}
    assert n["UTCHMMA"] >= 9 and n["LDTM"] >= 2 and n["UTCBAR"] >= 1 and n["FFMA2"] < 400 and n["UBLKCP"] >= 2, (key, n)
    mma = [l for l in body if "UTCHMMA" in l]
    assert all(len(re.findall(r"gdesc\[", l)) >= 2 for l in mma), mma[:2]


@pytest.mark.parametrize("key", [k for k in HOT if k not in HOT_TC_SS])       # en1 left the FMA pipe there: see the test above
def test_convolutions_are_packed_ffma2(sass, key):
    body = _find(sass, key)
    f2 = sum(1 for l in body if "FFMA2" in l)
    f1 = sum(1 for l in body if re.search(r"\sFFMA\s", l))
    assert f2 >= 100 and f2 >= 1.5 * f1, (key, f2, f1)


def test_default_kernels_do_not_spill():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([CUOBJDUMP, "-res-usage", B.build()], capture_output=True, text=True, check=True).stdout
    fn = None
    seen = 0
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1); continue
        m = re.search(r"STACK:(\d+)", line)
        if m and fn and any(k in fn for k in HOT):
            seen += 1
            assert int(m.group(1)) <= 16, (fn, line.strip())
    assert seen == len(HOT)
