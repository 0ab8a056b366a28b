"""Copies the reference's OWN unit tests for the hot path into tests/golden/ref_tests/ (test fixtures, category (b)): they are
executed -- unmodified -- against this repo's mirror of the API by tests/test_gpu_zzz_reference_suite.py on the GPU box, where
/root/reference does not exist.  Run in the build container: ``python tests/golden/make_ref_tests.py``.

Files (reference tests/checkpointing/unit/): __init__.py (TempNamedDir), conftest.py, test_utilities.py, test_async_save.py,
test_basic_local.py, test_cleanup.py, test_async_writer.py.  Nothing under nvidia-resiliency-ext_b200/ derives from them."""
import shutil
from pathlib import Path

SRC = Path("/root/reference/tests/checkpointing/unit")
DST = Path(__file__).resolve().parent / "ref_tests" / "tests" / "checkpointing" / "unit"
FILES = ["__init__.py", "conftest.py", "test_utilities.py", "test_async_save.py", "test_basic_local.py", "test_cleanup.py",
         "test_async_writer.py"]

if __name__ == "__main__":
    DST.mkdir(parents=True, exist_ok=True)
    for name in FILES:
        shutil.copyfile(SRC / name, DST / name)
    # test_async_writer.py imports `tests.checkpointing.unit` absolutely; an unrelated regular package called `tests` in
    # site-packages would shadow a namespace package here, so the two directories above `unit` become regular packages
    for pkg in (DST.parent, DST.parent.parent):
        (pkg / "__init__.py").touch()
    print(f"copied {len(FILES)} files to {DST}")
