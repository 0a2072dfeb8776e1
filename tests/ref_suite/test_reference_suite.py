"""Run the reference's own test-suite against this package.

`alias_plugin` makes `import lintsampler` resolve to `lintsampler_b200`; the reference's
`tests/` directory (never copied into this repo) is then collected and run unchanged in a child
pytest.  It needs BOTH a CUDA device (the drop-in has no CPU path; the reference's modules build
grids at import time) and a reference checkout (`$LINTSAMPLER_REF_TESTS`, default
`/root/reference/tests`), so it skips on the build container (no GPU) and on the GPU boxes of this
project (no checkout); it is for a maintainer's workstation:

    LINTSAMPLER_REF_TESTS=~/src/lintsampler/tests pytest tests/ref_suite -m gpu
"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_TESTS = os.environ.get("LINTSAMPLER_REF_TESTS", "/root/reference/tests")


def _child_env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT, HERE, env.get("PYTHONPATH", "")])
    return env


def test_alias_plugin_binds_the_drop_in():
    """CPU: under the plugin `import lintsampler` is this package, with the reference's public names."""
    code = ("import alias_plugin, lintsampler, lintsampler_b200 as lb; "
            "assert lintsampler is lb; "
            "from lintsampler import LintSampler, DensityGrid, DensityTree, DensityStructure")
    r = subprocess.run([sys.executable, "-c", code], env=_child_env(), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.gpu
def test_reference_suite_passes_against_the_drop_in(tmp_path):
    import torch
    if not os.path.isdir(REF_TESTS):
        pytest.skip(f"no reference checkout at {REF_TESTS}")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    r = subprocess.run([sys.executable, "-m", "pytest", REF_TESTS, "-p", "alias_plugin", "-q", "-x",
                        "--rootdir", str(tmp_path), "-p", "no:cacheprovider"],
                       env=_child_env(), cwd=str(tmp_path), capture_output=True, text=True, timeout=3600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
