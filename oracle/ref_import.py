"""Import the UNMODIFIED reference package from /root/reference.  TEST INFRASTRUCTURE ONLY.

Used by ``oracle/gen_golden.py`` (golden-vector generation) and by the ``not gpu`` tests that
re-check the restatement in ``oracle/ref_torch.py`` when /root/reference is present (it is not
on the GPU box; nothing that runs there may call this).

The reference has two module-level imports that are not installed in this image and are not
on the hot path: ``tensorboardX`` (colvarsfinder/core.py:50) and ``openmm``
(core.py:58, utils.py:57-58).  They are replaced by empty stand-ins before the import.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


class _NullWriter:
    """Stand-in for tensorboardX.SummaryWriter (core.py:143); swallows add_scalar calls."""

    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def close(self):
        pass


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "colvarsfinder", "core.py"))


def load():
    """Return the reference modules (core, nn, utils) imported under the name ``colvarsfinder_ref``."""
    if not available():
        raise RuntimeError("reference tree not present at /root/reference")
    if "colvarsfinder_ref" in sys.modules:
        m = sys.modules["colvarsfinder_ref"]
        return m.core, m.nn, m.utils
    for name in ("tensorboardX", "openmm", "openmm.app", "openmm.unit"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["tensorboardX"].SummaryWriter = _NullWriter
    sys.modules["openmm"].unit = sys.modules["openmm.unit"]
    sys.modules["openmm"].app = sys.modules["openmm.app"]
    # the reference does `from colvarsfinder.nn import ...` internally, so it has to be importable as
    # `colvarsfinder`; import it, then rename so it cannot shadow this repository's own package.
    saved = {k: v for k, v in sys.modules.items() if k == "colvarsfinder" or k.startswith("colvarsfinder.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        core = importlib.import_module("colvarsfinder.core")
        nn = importlib.import_module("colvarsfinder.nn")
        utils = importlib.import_module("colvarsfinder.utils")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    pkg = sys.modules["colvarsfinder"]
    for k in [k for k in sys.modules if k == "colvarsfinder" or k.startswith("colvarsfinder.")]:
        sys.modules[k.replace("colvarsfinder", "colvarsfinder_ref", 1)] = sys.modules.pop(k)
    sys.modules.update(saved)
    pkg.core, pkg.nn, pkg.utils = core, nn, utils
    return core, nn, utils


class FakeTrajectory:
    """Duck-typed stand-in for utils.WeightedTrajectory (the tasks only read
    .trajectory, .weights, .dt -- core.py:329,343-346)."""

    def __init__(self, trajectory, weights, dt=1.0):
        self.trajectory = trajectory
        self.weights = weights
        self.dt = dt
        self.n_frames = trajectory.shape[0]
