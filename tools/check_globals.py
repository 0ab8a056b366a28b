"""Poor man's linter (none is installed in the image): import every module of the package and report global names that a
function loads but that neither the module nor builtins define -- the kind of typo that only shows on a code path that needs a
GPU.  Usage: python tools/check_globals.py"""
import builtins
import dis
import importlib
import pkgutil
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "nvidia-resiliency-ext_b200"))
sys.path.insert(0, str(ROOT))


def code_objects(code):
    yield code
    for const in code.co_consts:
        if isinstance(const, types.CodeType):
            yield from code_objects(const)


def check_module(mod) -> list:
    problems = []
    src = getattr(mod, "__file__", None)
    if not src or not src.endswith(".py"):
        return problems
    top = compile(Path(src).read_text(), src, "exec")
    known = set(vars(mod)) | set(vars(builtins))
    for code in code_objects(top):
        for ins in dis.get_instructions(code):
            if ins.opname == "LOAD_GLOBAL" and ins.argval not in known:  # LOAD_NAME = class bodies / module level
                problems.append(f"{src}:{ins.positions.lineno if ins.positions else '?'}: {code.co_name} uses undefined global {ins.argval!r}")
    return problems


def check_abi_calls() -> list:
    """Every ``<something>.nvrx_xxx(...)`` call in the package must name an exported symbol and pass as many arguments as the
    ctypes signature declares (ctypes only checks that at call time -- on a GPU box)."""
    import ast

    from nvidia_resiliency_ext.checkpointing.b200 import _cabi

    lib = _cabi.lib()
    problems = []
    for path in list((ROOT / "nvidia-resiliency-ext_b200" / "nvidia_resiliency_ext").rglob("*.py")) + [ROOT / "bench.py", ROOT / "bench_c4.py", ROOT / "__graft_entry__.py"] + list((ROOT / "tools").glob("*.py")):
        tree = ast.parse(path.read_text())
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute) and node.attr.startswith("nvrx_") and node.attr not in _cabi.EXPORTED_SYMBOLS and node.attr != "nvrx_drain_aware":
                problems.append(f"{path}:{node.lineno}: {node.attr} is not an exported symbol")
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in _cabi.EXPORTED_SYMBOLS:
                want = len(getattr(lib, node.func.attr).argtypes or [])
                if not any(isinstance(a, ast.Starred) for a in node.args) and len(node.args) != want:
                    problems.append(f"{path}:{node.lineno}: {node.func.attr} called with {len(node.args)} arguments, signature has {want}")
    return problems


def main():
    import nvidia_resiliency_ext

    problems = []
    mods = [m.name for m in pkgutil.walk_packages(nvidia_resiliency_ext.__path__, "nvidia_resiliency_ext.")]
    extra = ["bench", "bench_c4", "__graft_entry__", "oracle.snapshot_oracle", "oracle.reference_port"]
    for name in mods + extra:
        try:
            mod = importlib.import_module(name)
        except Exception as exc:  # noqa: BLE001 - optional dependencies (lightning)
            print(f"skip {name}: {type(exc).__name__}: {exc}")
            continue
        problems += check_module(mod)
    # GPU-only measurement tools that only define functions at import time
    from importlib import util as _util

    for tool in ("d2h_ceiling", "restore_breakdown", "overlap_timeline", "bench_replicate", "ncu_summary"):
        spec = _util.spec_from_file_location(f"tools_{tool}", ROOT / "tools" / f"{tool}.py")
        mod = _util.module_from_spec(spec)
        try:
            spec.loader.exec_module(mod)
        except Exception as exc:  # noqa: BLE001
            problems.append(f"tools/{tool}.py does not import: {type(exc).__name__}: {exc}")
            continue
        problems += check_module(mod)
    problems += check_abi_calls()
    print("\n".join(problems) if problems else f"{len(mods) + len(extra)} modules: no undefined globals, C-ABI calls match their signatures")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
