"""Static checks of the built sm_100a library (cuobjdump -sass, no GPU): the tensor kernels are on the Blackwell paths
(UTCHMMA = tcgen05.mma, UTMALDG = TMA load, LDTM = tcgen05.ld, no mma.sync), the default GEMM instantiation has no
register-spill loads in its epilogue (the regression of round 2: folding split-K into it made it spill and 1.7x slower
at K = 320), and the cross-rank push stores through the NVSwitch multicast address.  Same parser as tools/sass_opcodes.py
(profiles/r02_sass_opcodes.txt)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "univst_b200", "libunivst_b200.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not on PATH")
    assert os.path.exists(LIB), "the library is built by tests/conftest.py"
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line) if cur is not None else None
        if m:
            full = m.group(1)
            cur[full.split(".")[0]] += 1          # the opcode ...
            if full.count("."):
                cur[full] += 1                    # ... and the opcode with its modifiers
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True, check=True).stdout.split("\n")
    out = {}
    for mangled, name in zip(counts, names):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("uv::", "").replace("(int)", "").replace("(bool)", "")
        out[short] = counts[mangled]
    return out


def _kernels(sass, prefix):
    ks = {k: v for k, v in sass.items() if k.startswith(prefix)}
    assert ks, f"no kernel named {prefix}* in the library: {sorted(sass)[:8]} ..."
    return ks


def test_gemm_kernel_is_tcgen05_tma_and_does_not_spill(sass):
    ks = _kernels(sass, "gemm_tc_kernel<")
    for name, c in ks.items():
        assert c["UTCHMMA"] > 0 and c["UTMALDG"] > 0 and c["LDTM"] > 0 and c["UTCBAR"] > 0, (name, dict(c))
        assert c["HMMA"] == 0, name
    default = [c for n, c in ks.items() if re.match(r"gemm_tc_kernel<1, (false|0)>", n)]
    assert len(default) == 1, sorted(ks)
    assert default[0]["LDL"] == 0, "the default GEMM instantiation reloads spilled registers"


def test_attention_kernels_are_tcgen05_tma(sass):
    for prefix in ("attention_tc_split_kernel<", "attention_tc_kernel<"):
        for name, c in _kernels(sass, prefix).items():
            assert c["UTCHMMA"] > 0 and c["UTMALDG"] > 0 and c["LDTM"] > 0 and c["MUFU"] > 0, (name, dict(c))
            assert c["HMMA"] == 0, name


def test_cross_rank_kernels_use_multicast_and_system_scope_flags(sass):
    # multimem.st.relaxed.sys is a 16-byte system-scope STG in SASS (the switch replicates it by address)
    for name, c in _kernels(sass, "xrank_push_kernel").items():
        assert c["STG.E.128.STRONG.SYS"] > 0, (name, "the multicast store is missing from the cross-rank push")
    # the synchronisation tail: release stores / acquire loads of the epoch flags at system scope
    for name, c in _kernels(sass, "xrank_barrier_kernel").items():
        sys_ops = [k for k in c if k.endswith(".SYS") and k.split(".")[0] in ("ST", "STG", "LD", "LDG")]
        assert any(k.startswith(("ST", "STG")) for k in sys_ops) and any(k.startswith(("LD", "LDG")) for k in sys_ops), (name, sys_ops)
