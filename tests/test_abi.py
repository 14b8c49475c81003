"""The C-ABI shared library: loads, exports exactly the symbols include/b200pose.h declares, and
the package has no CPU path (checked in a clean interpreter without the test conftest)."""
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200pose.h")).read()
    return re.findall(r"^(?:int|long long|void|const char\*)\s+(b200_\w+)\s*\(", text, flags=re.M)


def test_header_is_generated_from_optable():
    from pypose_b200._optable import lie_symbols, lm_symbols, scan_symbols
    declared = set(header_symbols())
    assert {s for s, *_ in lie_symbols()} <= declared
    assert {s for s, *_ in lm_symbols()} <= declared
    assert {s for s, *_ in scan_symbols()} <= declared


def test_library_loads_and_exports_every_declared_symbol():
    from pypose_b200 import _C
    lib = ctypes.CDLL(_C.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 138
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_library_has_no_unresolved_symbols():
    """dlopen with RTLD_NOW: a template declared in one translation unit and defined in another namespace would only fail
    at its first call on a GPU box."""
    from pypose_b200 import _C
    ctypes.CDLL(_C.LIB_PATH, mode=os.RTLD_NOW)


def test_library_exports_nothing_undeclared():
    from pypose_b200 import _C
    out = subprocess.run(["nm", "-D", "--defined-only", _C.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and "b200_" in l}
    assert exported == set(header_symbols())


def test_no_cpu_fallback_in_shipped_package():
    code = (
        "import torch, pypose_b200 as pp\n"
        "x = pp.randn_so3(4)\n"
        "try:\n"
        "    x.Exp()\n"
        "except NotImplementedError as e:\n"
        "    assert 'CPU' in str(e); print('RAISED')\n"
    )
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert "RAISED" in r.stdout, r.stdout + r.stderr


def test_package_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pypose_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_fused_launch_sites_pass_as_many_arguments_as_the_abi_declares():
    """Every `_launch("b200_lm_*", ref, [args...], n)` in optim/_fused.py is checked against the op table (ctypes would
    only notice a wrong count at call time, on a GPU box)."""
    import re
    from pypose_b200._optable import LM_OPS
    arity = {base: len(args) for base, args, _ in LM_OPS}
    src = open(os.path.join(ROOT, "pypose_b200", "optim", "_fused.py")).read()
    seen = 0
    for m in re.finditer(r'_launch\("(b200_lm_\w+)",\s*\w+,\s*\[(.*?)\],\s*[\w\.\[\]\(\) ]+\)', src, re.S):
        name, args = m.group(1), m.group(2)
        depth, n = 0, 1
        for ch in args:
            depth += ch in "([{"
            depth -= ch in ")]}"
            n += ch == "," and depth == 0
        n += 5 * args.count("*seg")        # seg = [Y4, poses, pidx, cseg, split, tpi]
        n += 6 * args.count("_NOCOMM")     # *(comm.pcg_args() if comm is not None else _NOCOMM): 7 arguments
        assert n == arity[name], (name, n, arity[name])
        seen += 1
    assert seen >= 30


def test_step_argument_block_mirror_has_the_c_layout():
    """optim/_lmstep.py PgoStepArgs mirrors `b200_pgo_step_args` field by field: sizes must agree."""
    from pypose_b200 import _C
    from pypose_b200.optim._lmstep import PgoStepArgs
    f = _C.lib().b200_pgo_step_args_size
    f.restype = ctypes.c_longlong
    assert ctypes.sizeof(PgoStepArgs) == f()
    assert PgoStepArgs.nodes.offset == 4 * 4 + 6 * 8 + 5 * 8 + 5 * 8 + 14 * 8
    from pypose_b200.optim._lmstep import BaStepArgs
    f = _C.lib().b200_ba_step_args_size
    f.restype = ctypes.c_longlong
    assert ctypes.sizeof(BaStepArgs) == f()
