"""Loads the package directory `x265-amod_b200/` (hyphenated, so not importable by name) as the
module `x265_amod_b200`, plus its synthetic-sequence generator."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "x265-amod_b200")


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[PKG_DIR] if path.endswith("__init__.py") else None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_pkg():
    return _load("x265_amod_b200", os.path.join(PKG_DIR, "__init__.py"))


def load_synth():
    return _load("x265_amod_b200.synth", os.path.join(PKG_DIR, "synth.py"))
