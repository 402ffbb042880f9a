"""Registers the on-disk package directory ``hr-viton_b200/`` under the importable
name ``hrviton_b200`` (a hyphen cannot appear in a Python module name).

Every root-level drop-in module (``networks.py``, ``network_generator.py``,
``sync_batchnorm``), ``bench.py``, ``__graft_entry__.py`` and the tests call
``hrv_loader.load()`` first and then ``import hrviton_b200.<sub>`` normally.
"""
import importlib.util
import os
import sys

_NAME = "hrviton_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "hr-viton_b200")


def load():
    mod = sys.modules.get(_NAME)
    if mod is not None:
        return mod
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
