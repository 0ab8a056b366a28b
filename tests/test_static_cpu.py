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
