"""The C-ABI library builds, loads, exports every symbol include/ira.h declares, its structs have
the layout the ctypes binding assumes, and it fails loudly without a GPU.  No compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ira.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ira_[a-z0-9_A-Z]+)\s*\(", src)))


def test_header_symbols_are_exported(built_lib):
    from irotavg_b200 import _lib
    lib = C.CDLL(built_lib)
    declared = _declared()
    assert declared, "no declarations parsed"
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in ira.h but not exported: {missing}"
    assert sorted(_lib.SYMBOLS) == declared


def test_struct_layout_matches_ctypes(tmp_path, built_lib):
    from irotavg_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ira.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    "sizeof(ira_options), sizeof(ira_stats), offsetof(ira_options, cg_rtol),"
                    "offsetof(ira_stats, score), offsetof(ira_stats, n_residual), sizeof(ira_mst_stats),"
                    "offsetof(ira_mst_stats, t_ms), offsetof(ira_options, shard_mode), offsetof(ira_options, pair_theta3));"
                    "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [int(v) for v in out] == [C.sizeof(_lib.Options), C.sizeof(_lib.Stats), _lib.Options.cg_rtol.offset,
                                     _lib.Stats.score.offset, _lib.Stats.n_residual.offset, C.sizeof(_lib.MstStats),
                                     _lib.MstStats.t_ms.offset, _lib.Options.shard_mode.offset,
                                     _lib.Options.pair_theta3.offset]


def test_header_is_plain_c(tmp_path):
    prog = tmp_path / "c.c"
    prog.write_text('#include "ira.h"\nint main(void){ira_options o; (void)o; return IRA_ABI_VERSION - 2;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c",
                    str(prog), "-o", str(tmp_path / "c.o")], check=True)


def test_defaults_and_status_strings(built_lib):
    from irotavg_b200 import _lib
    lib = _lib.load()
    assert lib.ira_abi_version() == 2
    o = _lib.Options()
    assert lib.ira_options_default(C.byref(o)) == 0
    assert o.device == -1 and o.world_size == 1 and o.cg_rtol > 0 and o.cg_max_iters > 0
    for s in range(9):
        assert lib.ira_status_string(s)
    assert b"Unknown cost" in lib.ira_status_string(5)


@pytest.mark.skipif(os.environ.get("IRA_EXPECT_GPU") == "1", reason="GPU box")
def test_fails_loudly_without_gpu(built_lib):
    import irotavg_b200 as ira
    if ira.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(ira.IraError) as ei:
        ira.Solver()
    assert ei.value.status == 2     # IRA_ERR_NO_DEVICE: no CPU fallback exists
    with pytest.raises(ira.IraError):
        ira.irls(np.zeros((1, 4)), [[0, 1]], None, ira.L1, 0.1, np.zeros((2, 4)), 1, 1, 1e-3)


def test_product_does_not_import_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "irotavg_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{fn} mentions oracle/"
    code = "import sys; import irotavg_b200; sys.exit(any(m.startswith('oracle') for m in sys.modules))"
    assert subprocess.run([sys.executable, "-c", code], cwd=ROOT).returncode == 0
