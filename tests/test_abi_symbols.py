"""The C-ABI library loads on a CPU-only box and exports every symbol include/qconsensus.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

from quoracle_b200 import native

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "qconsensus.h").read_text()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(qc_[a-z0-9_]+)\s*\(", body)))


def test_header_and_binding_list_agree():
    assert declared_functions() == sorted(native.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = native.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_status_strings_and_atoms():
    lib = native.load_library()
    assert lib.qc_abi_version() == 1
    atoms = {0: "ok", -1: "invalid_input", -2: "not_configured", -3: "service_unavailable", -4: "service_unavailable",
             -5: "service_unavailable", -6: "context_length_exceeded", -7: "bad_request", -8: "invalid_input",
             -9: "service_unavailable"}
    for code, atom in atoms.items():
        assert lib.qc_status_atom(code).decode() == atom
        assert lib.qc_strerror(code).decode() != "unknown status"
    assert lib.qc_strerror(-99).decode() == "unknown status"


def test_no_cpu_fallback_without_a_gpu():
    """On a box without an sm_100 device the context cannot be created: the product has no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; covered by the -m gpu tests")
    with pytest.raises(native.QcError) as e:
        native.Context(0)
    assert e.value.status == -3 and e.value.atom == "service_unavailable"


def test_null_arguments_are_rejected_not_dereferenced():
    lib = native.load_library()
    assert lib.qc_ctx_create(None, 0, None) == -1
    assert lib.qc_ctx_synchronize(None) == -1
    assert lib.qc_similarity_batch(None, None, 1, 2, 3, None, None, None, None) == -1
    assert lib.qc_launch_count(None) == 0
    lib.qc_ctx_destroy(None)


def test_c_example_compiles_and_links_against_the_header(tmp_path):
    """examples/c_abi_decide.c is plain C against include/qconsensus.h: it must build with -Wall -Werror and link
    against the in-tree library without any Python or torch in sight (no GPU needed to link)."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    lib = root / "quoracle_b200" / "lib"
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-std=c11", f"-I{root / 'include'}", str(root / "examples" / "c_abi_decide.c"),
                    f"-L{lib}", "-lqconsensus", f"-Wl,-rpath,{lib}", "-o", str(tmp_path / "c_abi_decide")], check=True)


def test_hot_kernels_are_blackwell_native_sass():
    """Static check of the built library (cuobjdump, no GPU): the encoder and the large-pool similarity kernel are
    tcgen05 / TMEM / TMA code -- UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / .st), UTMALDG (TMA, with GATHER4 in
    the encoder), UTCBAR (tcgen05.commit) -- and nothing in the binary comes from a library (profiles/sass_census_r2.md
    is the same count, per kernel)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    from quoracle_b200 import build
    native.load_library()                                    # fails loudly, like every other test, when the library was not built
    sass = subprocess.run([cuobjdump, "-sass", str(build.LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
        elif cur is not None:
            kernels[cur].append(line)
    enc = {k: "\n".join(v) for k, v in kernels.items() if "qc_encode_kernel" in k}
    tc = {k: "\n".join(v) for k, v in kernels.items() if "qc_sim_tc_kernel" in k}
    assert len(enc) >= 30 and len(tc) == 3
    for name, body in enc.items():
        for op in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "MUFU.TANH"):
            assert op in body, (name, op)
    assert any("GATHER4" in body for body in enc.values())
    for name, body in tc.items():
        for op in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR"):
            assert op in body, (name, op)
    assert not any(re.search(r"cutlass|cublas|cudnn|triton", k, re.I) for k in kernels)
