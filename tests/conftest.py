import functools
import inspect
import json
import os
import subprocess
import sys
import tempfile

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "log-neural-cdes_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Build (idempotent) and load the C-ABI library."""
    from lncde import _build, _lib

    _build.build_library()
    return _lib.lib()


def rel_err(a, b):
    """Norm-relative error per tensor: max|a-b| / max|b|."""
    import numpy as np

    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    err = float(np.abs(a - b).max() / (den if den > 0 else 1.0))
    _log_err(err, a.shape)
    return err


def _log_err(err, shape):
    """Measured parity errors of a run, one JSON line per comparison (LNCDE_ERR_LOG=<file>): the tolerances asserted in
    the GPU tests are set from these records (profiles/r02_parity_errors.md)."""
    log = os.environ.get("LNCDE_ERR_LOG")
    if not log:
        return
    import traceback

    fr = next((f for f in reversed(traceback.extract_stack()[:-2]) if "/tests/" in f.filename), None)
    with open(log, "a") as fh:
        fh.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", "").rsplit(" (", 1)[0], "err": err,
                             "shape": list(shape), "at": f"{os.path.basename(fr.filename)}:{fr.lineno}" if fr else ""}) + "\n")


# ---- child-process isolation for GPU tests of kernels that have not been on the hardware yet -----------------
# A kernel that faults leaves a sticky CUDA error behind and one that hangs never returns: either would take
# every later test of the session with it.  Tests decorated with @isolated_gpu therefore run in ONE child pytest
# process per test file; the child streams a verdict per test (hook below), the parent only reports them.
INNER_ENV = "LNCDE_TEST_INNER"
RESULT_LOG_ENV = "LNCDE_TEST_RESULT_LOG"
INNER = bool(os.environ.get(INNER_ENV))
_inner_cache = {}


def pytest_runtest_logreport(report):
    log = os.environ.get(RESULT_LOG_ENV)
    if not log or not (report.when == "call" or (report.when == "setup" and report.outcome != "passed")):
        return
    with open(log, "a") as fh:
        fh.write(json.dumps({"id": report.nodeid.split("::", 1)[1], "outcome": report.outcome,
                             "text": "" if report.passed else str(report.longrepr)[-4000:]}) + "\n")


def _inner_results(path, timeout_s):
    if path not in _inner_cache:
        res, note = {}, ""
        with tempfile.TemporaryDirectory() as tmp:
            log = os.path.join(tmp, "verdicts.jsonl")
            env = dict(os.environ, **{INNER_ENV: "1", RESULT_LOG_ENV: log})
            cmd = [sys.executable, "-m", "pytest", os.path.relpath(path, REPO), "-q", "-m", "gpu", "-p", "no:cacheprovider"]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env, cwd=REPO)
                note = f"child pytest exit {r.returncode}: {(r.stdout + r.stderr)[-1500:]}"
            except subprocess.TimeoutExpired:
                note = f"child pytest killed after {timeout_s} s (a kernel of this file hung)"
            if os.path.exists(log):
                with open(log) as fh:
                    for line in fh:
                        rec = json.loads(line)
                        res[rec["id"]] = (rec["outcome"], rec["text"])
        _inner_cache[path] = (res, note)
    return _inner_cache[path]


def isolated_gpu(fn=None, *, timeout_s=900):
    """Run the decorated GPU test inside the file's child pytest process (see above) and report its verdict here."""
    if fn is None:
        return functools.partial(isolated_gpu, timeout_s=timeout_s)
    path = os.path.abspath(inspect.getsourcefile(fn))

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if INNER:
            return fn(*args, **kwargs)
        name = os.environ["PYTEST_CURRENT_TEST"].rsplit(" (", 1)[0].split("::", 1)[1]
        res, note = _inner_results(path, timeout_s)
        if name not in res:
            pytest.fail(f"no verdict for {name}: {note}", pytrace=False)
        outcome, text = res[name]
        if outcome == "skipped":
            pytest.skip(text)
        if outcome != "passed":
            pytest.fail(text, pytrace=False)

    return wrapper
