"""The built library really contains Blackwell-native code (B200_PROFILING.md "What proves a Blackwell-native kernel"):
tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA bulk copies -> UBLKCP, for sm_100a only.  CPU-only: reads the SASS with cuobjdump."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "totsu_b200", "libtotsu_b200.so")


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_library_sass_has_tcgen05_tmem_and_tma():
    assert os.path.exists(LIB), "run __graft_entry__.build() first"
    elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in elf and all("sm_100a" in ln for ln in elf.splitlines() if ln.strip().startswith("ELF file")), elf
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    funcs = {}
    cur = None
    for ln in sass.splitlines():
        if "Function :" in ln:
            cur = ln.split("Function :")[1].strip()
            funcs[cur] = set()
        elif cur is not None:
            for m in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "UCGABAR", "SYNCS"):
                if m in ln:
                    funcs[cur].add(m)
    tc = [f for f in funcs if "symm_gemm_tc_kernel" in f]
    assert tc and all({"UTCHMMA", "LDTM", "UTCBAR"} <= funcs[f] for f in tc), {f: funcs[f] for f in tc}       # tcgen05.mma / ld / commit
    assert any("UCGABAR" in funcs[f] for f in tc)                                                           # cluster barrier of the split-K exchange
    stream = [f for f in funcs if "stream_kernel" in f]
    assert stream and all({"UBLKCP", "SYNCS"} <= funcs[f] for f in stream), {f: funcs[f] for f in stream}     # TMA bulk copies on mbarriers
    assert any("vprog_kernel" in f and "UCGABAR" in funcs[f] for f in funcs)
    assert re.search(r"(?<![A-Z])(HMMA|HGMMA|QGMMA|IGMMA)\b", sass) is None                                   # no legacy / Hopper tensor path
