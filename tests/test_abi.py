"""CPU: the C-ABI library builds, loads and exports every symbol include/l2i_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "l2i_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(l2i_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from latent2im_b200 import _native as nt
    if not os.path.exists(nt.lib_path()):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(nt.lib_path())
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in l2i_b200.h but not exported"
    assert set(declared) == set(nt.EXPORTED_SYMBOLS), set(declared) ^ set(nt.EXPORTED_SYMBOLS)
    lib.l2i_abi_version.restype = ctypes.c_int
    assert lib.l2i_abi_version() == 1


def test_missing_cuda_fails_loudly():
    """No CPU fallback: CPU tensors raise exactly like the reference's CHECK_CUDA."""
    import torch
    from latent2im_b200.graphs.stylegan_v2_real.op import fused_leaky_relu, upfirdn2d
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        fused_leaky_relu(torch.zeros(1, 2, 3, 3), torch.zeros(2))
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        upfirdn2d(torch.zeros(1, 2, 3, 3), torch.ones(1, 1))


def test_state_dict_layout_matches_reference():
    import json
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_state_dict_keys.json")))
    for size, keys in ref.items():
        g = Generator(int(size), 512, 8)
        mine = [[k, list(v.shape)] for k, v in g.state_dict().items()]
        assert mine == keys


def test_weight_version_check_sees_every_kind_of_update():
    """The per-call weight-version check reads (data_ptr, version) through cached (module, name) slots instead of walking
    nn.Module.parameters(): in-place updates, load_state_dict and a Parameter assigned to an existing module must all change
    the fingerprint; invalidate_native() forgets the cache (sub-module surgery)."""
    import torch
    from latent2im_b200.graphs.stylegan_v2_real.networks import Generator
    gen = Generator(16, 32, 2)

    def fingerprint():
        return tuple((p.data_ptr(), p._version) for p in gen._param_list())

    assert len(gen._param_list()) == len(list(gen.parameters()))
    assert {id(p) for p in gen._param_list()} == {id(p) for p in gen.parameters()}
    f0 = fingerprint()
    assert fingerprint() == f0
    with torch.no_grad():
        gen.conv1.conv.weight.mul_(1.5)                       # in-place update: version counter
    f1 = fingerprint()
    assert f1 != f0
    gen.load_state_dict({k: v.clone() for k, v in gen.state_dict().items()})   # copy_ into the same storage: version counter
    f2 = fingerprint()
    assert f2 != f1
    gen.conv1.conv.weight = torch.nn.Parameter(gen.conv1.conv.weight.detach().clone())   # new Parameter on an existing module
    f3 = fingerprint()
    assert f3 != f2 and {id(p) for p in gen._param_list()} == {id(p) for p in gen.parameters()}
    gen.invalidate_native()
    assert "_param_slots" not in gen.__dict__ and fingerprint() == f3
