"""No linter ships with the image, and most of the engine only runs on a GPU box: at least make sure that no function of the
package, bench.py or __graft_entry__.py loads a global name that does not exist (tools/check_globals.py)."""
import importlib.util

from conftest import ROOT


def test_no_undefined_globals(capsys):
    spec = importlib.util.spec_from_file_location("check_globals", ROOT / "tools" / "check_globals.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rc = mod.main()
    out = capsys.readouterr().out
    assert rc == 0, out


def test_the_product_never_imports_the_oracle_or_the_test_helpers():
    """The oracle is the checker, not a fallback: nothing under the package (or in the C sources) may import or name it."""
    import re

    pkg = ROOT / "nvidia-resiliency-ext_b200"
    offenders = []
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list((ROOT / "include").glob("*.h")):
        text = path.read_text()
        if re.search(r"^\s*(from|import)\s+(oracle|_fake_device|tests)\b", text, re.M) or re.search(r"#\s*include[^\n]*oracle", text):
            offenders.append(str(path))
    assert not offenders, offenders


def test_every_environment_switch_is_documented():
    """INTEGRATION.md's table is where users look: every NVRX_B200_* name the package or the C sources read must be in it
    (families like NVRX_B200_*_CTAS_PER_SM count through their wildcard row)."""
    import re

    pkg = ROOT / "nvidia-resiliency-ext_b200"
    names = set()
    for path in [p for ext in ("*.py", "*.cu", "*.cuh", "*.h", "*.cpp") for p in pkg.rglob(ext)]:
        names.update(re.findall(r"NVRX_B200_[A-Z0-9_]+", path.read_text()))
    doc = (ROOT / "INTEGRATION.md").read_text()
    wildcards = [re.compile(re.escape(w).replace(r"\*", "[A-Z0-9_]+") + r"$") for w in re.findall(r"NVRX_B200_[A-Z0-9_*]*\*[A-Z0-9_*]*", doc)]
    missing = sorted(n for n in names if n not in doc and not any(w.match(n) for w in wildcards))
    assert not missing, missing
