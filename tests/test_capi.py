"""The C-ABI shared library loads on a CPU-only box and exports every symbol the header declares."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from circuitree_b200 import _capi, _capi_defaults

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "circuitree_b200.h"


def declared_symbols():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ct_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(str(_capi.library_path()))
    names = declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in {HEADER.name} but not exported"


def test_binding_covers_header():
    from circuitree_b200 import native

    lib = native._lib()            # binds the ct_tree_* half on top of _capi.lib()
    unbound_ok = ("ct_version", "ct_last_error", "ct_tree_last_error")
    for name in declared_symbols():
        assert getattr(lib, name).argtypes is not None or name in unbound_ok, name


def test_default_config_matches_spec_constants():
    cfg = _capi.default_config()
    d = _capi_defaults.SPEC_DEFAULTS
    assert (cfg.nt, cfg.decim) == (d["nt"], d["decim"])
    assert cfg.dt == d["dt"] and cfg.init_mean == d["init_mean"]
    assert np.allclose(cfg.ranges_array(), np.array(d["ranges"]), atol=1e-6)
    assert np.allclose(cfg.ranges_array(12), np.array(d["ranges"] + d["dimer_ranges"]), atol=1e-6)
    assert cfg.n_dec == 125 <= _capi.MAX_NDEC


def test_compile_topology_validation_without_gpu():
    h = _capi.compile_topology(3, [], [(0, 1), (1, 2), (2, 0)])
    assert (h >> 56) & 15 == 3
    # edge (i, j) sits at bits 2*(5*i + j): inhibition = 2
    assert (h >> (2 * (0 * 5 + 1))) & 3 == 2 and (h >> (2 * (2 * 5 + 0))) & 3 == 2
    assert _capi.compile_topology(2, [(0, 1)], [(1, 0)]) != _capi.compile_topology(2, [(1, 0)], [(0, 1)])
    with pytest.raises(_capi.EngineError):
        _capi.compile_topology(3, [(0, 3)], [])
    with pytest.raises(_capi.EngineError):
        _capi.compile_topology(3, [(0, 1)], [(0, 1)])          # duplicate edge
    with pytest.raises(_capi.EngineError):
        _capi.compile_topology(6, [], [])                       # too many species for the GPU path


def test_compile_dimer_topology_without_gpu():
    """k = 3 rows (monomer1 <= monomer2, target) as DimersGrammar.parse_genotype returns them
    (reference circuitree/models.py:844-884)."""
    from circuitree_b200 import DimersGrammar

    g = DimersGrammar(components=["A", "B", "C"], regulators=["R"], interactions=["activates", "inhibits"])
    comps, regs, act, inh = g.parse_genotype("*ABC+R::ARaB_BBiC_CCiA")
    h = _capi.compile_topology(len(comps) + len(regs), act, inh, k=3)
    assert _capi.is_dimer_handle(h) and (h >> 56) & 15 == 4
    assert _capi.compile_topology(4, act, inh, k=3) == h                  # equal topologies share a handle
    assert _capi.compile_topology(4, inh, act, k=3) != h
    assert not _capi.is_dimer_handle(_capi.compile_topology(3, [], [(0, 1)]))
    with pytest.raises(_capi.EngineError):
        _capi.compile_topology(3, [(1, 0, 2)], [], k=3)         # monomers must be sorted
    with pytest.raises(_capi.EngineError):
        _capi.compile_topology(3, [(0, 3, 2)], [], k=3)         # monomer out of range
    with pytest.raises(_capi.EngineError):
        _capi.compile_topology(3, [(0, 1, 2)], [(0, 1, 2)], k=3)   # duplicate interaction
    with pytest.raises(_capi.EngineError):
        _capi.compile_topology(3, [(0, 0, 0)] * 11, [], k=3)     # more than CT_MAX_INTERACTIONS


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to exist instead of computing on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(_capi.EngineError, match="no CUDA device|CUDA"):
        _capi.Engine(0)


def test_product_never_imports_oracle():
    for path in (ROOT / "circuitree_b200").rglob("*.py"):
        src = path.read_text()
        assert "oracle" not in src.replace("no oracle", ""), f"{path} mentions the oracle"
    for path in (ROOT / "circuitree_b200" / "csrc").glob("*"):
        assert "oracle" not in path.read_text(), path


def test_fast_loop_instruction_budget():
    """The headline kernel is bound by instruction issue: its roofline fraction is
    I_step / (instructions executed per step) x issue efficiency, so the instruction count of the main path of the
    built library is part of the contract.  Static count from the SASS (scripts/sass_hot_loop.py,
    profiles/r02_final_fast_loop_sass.txt): 280 per Philox block = 140 per step against I_step = 113.5."""
    import shutil
    import subprocess
    import sys

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "sass_hot_loop.py")], capture_output=True, text=True, check=True).stdout
    head = out.splitlines()[1]
    n = int(head.split("point): ")[1].split()[0])
    assert 227 <= n <= 290, head          # 227 = 2 x I_step: fewer would mean the walk lost the path
    assert "MUFU 4" in out.splitlines()[3] and "IMAD 21" in out.splitlines()[3]      # 2 x (lg2 + rcp), Philox + 1


def test_explicit_parameters_must_be_finite_and_non_negative():
    """A NaN, infinite or negative rate would stall a trajectory's clock (a kernel that never ends): the host-buffer
    entry points refuse such parameter vectors with CT_ERR_INVALID; the check itself needs no GPU."""
    rg = np.random.default_rng(0)
    ok = rg.random((1 << 18, 10), dtype=np.float32)            # 2.6 M values: the multi-threaded scan
    _capi.check_params(ok)
    _capi.check_params(np.zeros((0, 10), dtype=np.float32))
    for bad_value in (np.nan, np.inf, -1.0e-3):
        for where in ((0, 0), (7, 3), ((1 << 18) - 1, 9), (200_000, 5)):
            bad = ok.copy()
            bad[where] = bad_value
            with pytest.raises(_capi.EngineError, match=f"parameter {where[1]} of trajectory {where[0]} "):
                _capi.check_params(bad)
    small = np.ones((3, 12), dtype=np.float32)
    small[2, 11] = np.nan
    with pytest.raises(_capi.EngineError, match="parameter 11 of trajectory 2 "):
        _capi.check_params(small)
