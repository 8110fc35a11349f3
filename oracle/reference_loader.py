"""Import the live reference package (build container only) -- TEST INFRASTRUCTURE ONLY.

``/root/reference`` exists in the build container and NOT on the GPU box, so this module
is used only by ``oracle/make_golden.py`` (to generate the committed fixtures) and by the
``not gpu`` tests that re-pin the oracle when the reference tree happens to be present.

The reference's ``torchcde/__init__.py:7`` imports ``solver.py``, whose first lines import
``torchdiffeq`` and ``torchsde`` (solver.py:2-3).  Neither is installed.  We therefore put
two stand-in modules in ``sys.modules`` before importing:
  * ``torchdiffeq`` -> ``oracle.odeint_port`` (the restated fixed-grid solver), so that the
    reference's own ``cdeint`` / ``_check_compatability`` / ``_VectorField`` / output permute
    run unmodified around it;
  * ``torchsde``    -> an empty module (that backend is out of scope).
Nothing is copied from the reference tree; it is imported where it lies.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TORCHCDE_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "torchcde", "__init__.py"))


def load_reference():
    """Return the reference ``torchcde`` module, or raise ``ImportError`` if the tree is absent."""
    if not reference_available():
        raise ImportError("reference tree not present at {}".format(REFERENCE_ROOT))
    if "torchcde" in sys.modules and getattr(sys.modules["torchcde"], "_b200_oracle_loaded", False):
        return sys.modules["torchcde"]
    from . import odeint_port

    diffeq = types.ModuleType("torchdiffeq")
    diffeq.odeint = odeint_port.odeint
    diffeq.odeint_adjoint = odeint_port.odeint_adjoint
    diffeq.__version__ = "port-of-0.2.x"
    sys.modules["torchdiffeq"] = diffeq
    sys.modules.setdefault("torchsde", types.ModuleType("torchsde"))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    module = importlib.import_module("torchcde")
    module._b200_oracle_loaded = True
    return module
