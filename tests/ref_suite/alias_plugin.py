"""pytest plugin: `import lintsampler` resolves to lintsampler_b200 (the drop-in under test).

Loaded with `-p tests.ref_suite.alias_plugin` by test_reference_suite.py before the reference's
own test modules are imported, so their `from lintsampler import LintSampler, DensityGrid, ...`
binds this package's classes.
"""
import sys

import lintsampler_b200

sys.modules["lintsampler"] = lintsampler_b200
for _name in ("DensityStructure", "DensityGrid", "DensityTree", "LintSampler"):
    assert hasattr(lintsampler_b200, _name), _name
