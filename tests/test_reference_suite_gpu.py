"""The reference's OWN pytest suite (pytorch_bayesian tests/, 34 tests: API contracts of every class, known-answer
contractions at 1e-5, KL > 0, prune fractions) run unmodified against the drop-in on the GPU, through the documented
alias `sys.modules['pytorch_bayesian'] = bayesianneuralnetworks_b200`.  The suite lives in the git-ignored
baseline/_ref/ (installed by baseline/install_ref.py, shipped to the GPU box with the snapshot)."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def test_reference_test_suite_passes_on_the_drop_in():
    if not os.path.isdir(os.path.join(REF, "tests")):
        pytest.skip("baseline/_ref not installed (run python baseline/install_ref.py where /root/reference exists)")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests") + os.pathsep + ROOT)
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "_reference_alias_plugin", "-p", "no:cacheprovider",
                          "--rootdir", REF, "-c", os.devnull, os.path.join(REF, "tests")],
                         cwd=REF, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    tail = "\n".join(res.stdout.strip().splitlines()[-60:])
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "reference_suite.log"), "w") as fh:
            fh.write(res.stdout)
    m = re.search(r"(\d+) passed", res.stdout)
    assert res.returncode == 0 and m and int(m.group(1)) >= 34, tail
