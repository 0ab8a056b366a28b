"""The torchrun runner of the reference's own unit tests (tests/test_gpu_zzz_reference_suite.py) itself, on CPU: results come back
through files, a process a test leaves behind neither blocks the call nor survives it, a time limit ends in an assertion."""
import os
import time

import pytest

import test_gpu_zzz_reference_suite as rs


def _fixture_tree(root, body):
    unit = root / "tests" / "checkpointing" / "unit"
    unit.mkdir(parents=True)
    (unit / "test_x.py").write_text(body)
    return root


LEAVES_A_PROCESS = """
import os, subprocess, sys
def test_leaves_a_process_behind():
    p = subprocess.Popen([sys.executable, "-c", "import time; time.sleep(600)"], start_new_session=True)   # inherits stdout / stderr
    open(os.environ["PID_FILE"], "w").write(str(p.pid))
def test_second():
    print("second ran")
"""


def _alive(pid):
    try:
        return open(f"/proc/{pid}/stat").read().split(")")[-1].split()[0] != "Z"
    except OSError:
        return False


def test_a_left_over_process_neither_blocks_nor_survives(tmp_path, monkeypatch):
    monkeypatch.setattr(rs, "FIXTURES", _fixture_tree(tmp_path / "fx", LEAVES_A_PROCESS))
    pid_file = tmp_path / "pid"
    monkeypatch.setenv("PID_FILE", str(pid_file))
    t0 = time.perf_counter()
    out = rs.run_reference_tests(["test_x.py"], 1, timeout=240)
    assert "2 passed" in out
    assert time.perf_counter() - t0 < 200  # the sleeper holds the inherited descriptors for 600 s
    pid = int(pid_file.read_text())
    deadline = time.time() + 10
    while _alive(pid) and time.time() < deadline:
        time.sleep(0.1)
    assert not _alive(pid)


def test_failures_and_time_limits_end_in_assertions_with_the_output(tmp_path, monkeypatch):
    monkeypatch.setattr(rs, "FIXTURES", _fixture_tree(tmp_path / "fx", "def test_no():\n    assert 1 == 2, 'marker-of-the-failure'\n"))
    with pytest.raises(AssertionError, match="marker-of-the-failure"):
        rs.run_reference_tests(["test_x.py"], 1, timeout=240)
    (tmp_path / "fx" / "tests" / "checkpointing" / "unit" / "test_x.py").write_text(
        "import time\ndef test_hangs():\n    print('before-the-hang', flush=True)\n    time.sleep(600)\n")
    t0 = time.perf_counter()
    with pytest.raises(AssertionError, match="no result within"):
        rs.run_reference_tests(["test_x.py"], 1, extra=["-s"], timeout=12)
    assert time.perf_counter() - t0 < 120


def test_group_counts_are_parsed_from_the_summary_line():
    assert rs.passed_count("....\n4 passed, 9 deselected, 15 warnings in 66.94s (0:01:06)\n") == 4
    assert rs.passed_count("1 passed in 0.5s") == 1
