"""The NIF shim (quoracle_b200/nif/qconsensus_nif.c) against the mock Erlang runtime -- the parts that need no GPU:
the module/export table a VM would see, its agreement with the stubs in elixir/quoracle_native.ex, badarg on
malformed terms, and the `{:error, atom}` a machine without a usable GPU gets (never a crash, never a raise)."""
import re
from pathlib import Path

import pytest

from tests import mock_erts
from tests.mock_erts import Atom, BadArg

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def vm():
    return mock_erts.MockVM()


def test_module_and_export_table(vm):
    assert vm.module() == "Elixir.Quoracle.Native"
    ex = vm.exports()
    for name, arity in [("ctx_create", 5), ("decide_batch", 8), ("embed_batch", 5), ("embed_ragged", 3),
                        ("decide_ragged", 6), ("similarity_batch", 6), ("first_match_batch", 7), ("chunk_mean", 5),
                        ("decide_batch_submit", 9), ("decide_ragged_submit", 7), ("decide_batch_wait", 2),
                        ("lesson_cache_create", 3), ("lesson_cache_put", 3), ("lesson_cache_gather", 2),
                        ("lesson_cache_match", 4)]:
        assert (name, arity) in ex, (name, arity)
        assert ex[(name, arity)] == 1, f"{name}/{arity} must run on a dirty CPU scheduler"
    assert ex[("launch_count", 1)] == 0


def test_elixir_stubs_match_the_nif_table(vm):
    """Every `def f(args), do: :erlang.nif_error(:nif_not_loaded)` stub of Quoracle.Native is exported by the NIF with
    the same arity, and vice versa (a mismatch makes :erlang.load_nif fail in a real VM)."""
    src = (ROOT / "quoracle_b200" / "elixir" / "quoracle_native.ex").read_text()
    stubs = set()
    for m in re.finditer(r"^\s*def (\w+)\(([^)]*)\), do: :erlang\.nif_error\(:nif_not_loaded\)", src, re.M):
        args = [a for a in m.group(2).split(",") if a.strip()]
        stubs.add((m.group(1), len(args)))
    assert stubs == set(vm.exports()), (stubs ^ set(vm.exports()))


def test_badarg_and_errors_without_touching_a_gpu(vm):
    with pytest.raises(BadArg):
        vm.call("ctx_create", b"zero", 0, 0, None, None)          # device id must be an integer
    with pytest.raises(BadArg):
        vm.call("decide_batch", Atom("not_a_ctx"), b"", b"", 0, 1, 1, b"", 0)
    with pytest.raises(BadArg):
        vm.call("decide_batch_wait", 7, 0)
    with pytest.raises(AttributeError):
        vm.call("decide_batch", 1, 2)                              # wrong arity: undefined function in a VM
    # weights of the wrong size are a plain error tuple
    assert vm.call("ctx_create", 0, 16, 384, b"\0" * 10, b"\0" * 10) == (Atom("error"), Atom("invalid_input"))


def test_no_device_is_an_error_tuple(vm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    assert vm.call("ctx_create", 0, 0, 0, None, None) == (Atom("error"), Atom("service_unavailable"))
    assert vm.live_resources() == 0
