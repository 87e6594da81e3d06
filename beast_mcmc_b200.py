"""Loader shim: makes the hyphenated package directory ``beast-mcmc_b200/`` importable as
``beast_mcmc_b200`` (a hyphen is not legal in a Python module name)."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "beast-mcmc_b200")
_spec = importlib.util.spec_from_file_location(
    "beast_mcmc_b200", os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["beast_mcmc_b200"] = _mod
_spec.loader.exec_module(_mod)
