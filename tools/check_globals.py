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


def main():
    import nvidia_resiliency_ext

    problems = []
    mods = [m.name for m in pkgutil.walk_packages(nvidia_resiliency_ext.__path__, "nvidia_resiliency_ext.")]
    for name in mods + ["bench", "__graft_entry__"]:
        try:
            mod = importlib.import_module(name)
        except Exception as exc:  # noqa: BLE001 - optional dependencies (lightning)
            print(f"skip {name}: {type(exc).__name__}: {exc}")
            continue
        problems += check_module(mod)
    print("\n".join(problems) if problems else f"{len(mods) + 2} modules: no undefined globals")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
