"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the
header declares, mirrors the config struct, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from after_b200 import build, _lib
    build.build()
    return _lib.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "after_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(after_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from after_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"libafter_b200.so does not export {s}"
        assert s in _lib.PROTOTYPES, f"python binding lacks a prototype for {s}"
    assert sorted(_lib.PROTOTYPES) == syms


def test_abi_version_and_build_info(lib):
    from after_b200 import _lib
    assert lib.after_abi_version() == _lib.ABI_VERSION == 3
    assert b"sm_100a" in lib.after_build_info()


def test_config_struct_layout_matches_header(lib, tmp_path):
    """Compile a two-line C program against the header and compare sizeof/offsets with the ctypes mirror."""
    import subprocess
    from after_b200 import _lib
    src = tmp_path / "sz.c"
    fields = ["abi_version", "drop_value", "max_steps", "ae_multipliers", "ae_max_samples", "se_in_size", "se_use_tanh",
              "max_cache_size", "un_in_size", "un_ratios", "un_use_res_last"]
    body = "".join(f'printf("%zu\\n", offsetof(after_config, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "after_b200.h"\n'
                   f'int main(void){{printf("%zu\\n", sizeof(after_config));{body}return 0;}}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    nums = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert nums[0] == C.sizeof(_lib.AfterConfig)
    for f, off in zip(fields, nums[1:]):
        assert getattr(_lib.AfterConfig, f).offset == off, f


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure(lib):
    from after_b200 import _lib
    cfg = _lib.AfterConfig()
    cfg.abi_version = _lib.ABI_VERSION
    h = C.c_void_p()
    rc = lib.after_create(C.byref(cfg), 0, C.byref(h))
    assert rc < 0
    assert b"no CPU fallback" in lib.after_last_error(None) or b"CUDA" in lib.after_last_error(None)
    from after_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine()


def test_plain_c_host_compiles_links_and_fails_loudly_without_gpu(lib, tmp_path):
    """examples/c_host.c: a C program binds the ABI with nothing but the header; without a CUDA device it must stop at
    after_create with the library's message (exit code 2), with one it samples (exit code 0)."""
    import subprocess
    libdir = os.path.join(ROOT, "after_b200", "lib")
    exe = tmp_path / "c_host"
    subprocess.check_call(["gcc", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_host.c"),
                           "-o", str(exe), "-L", libdir, "-lafter_b200", f"-Wl,-rpath,{libdir}"])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
        assert "sample ok" in r.stdout and "yes" in r.stdout
    else:
        assert r.returncode == 2, r.stdout + r.stderr
        assert "after_create failed" in r.stdout and ("no CPU fallback" in r.stdout or "CUDA" in r.stdout)
