"""Import the UNMODIFIED reference (under oracle/gym_stub) -- TEST INFRASTRUCTURE ONLY.

Search order: ``/root/reference`` (build container), then ``oracle/_ref`` (byte-compiled by
``oracle/build_ref.py``; the only form that reaches the GPU box).  ``available()`` says whether
either exists; nothing in the GPU tests or smoke() requires it.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUB = os.path.join(_HERE, "gym_stub")
_CANDIDATES = ["/root/reference", os.path.join(_HERE, "_ref")]
_REF_PACKAGES = ("envs", "rewards", "policies", "experiments", "training", "evaluation")


BC_SUFFIX = ".refbc"


def reference_root():
    if os.path.isfile(os.path.join(_CANDIDATES[0], "envs", "__init__.py")):
        return _CANDIDATES[0]
    if os.path.isfile(os.path.join(_CANDIDATES[1], "envs", "__init__" + BC_SUFFIX)):
        return _CANDIDATES[1]
    return None


class _ByteCodeFinder(importlib.abc.MetaPathFinder):
    """Resolves the reference's packages from oracle/_ref/**/*.refbc (sourceless byte code)."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] not in _REF_PACKAGES:
            return None
        for entry in ([self.root] if path is None else list(path)):
            finder = importlib.machinery.FileFinder(entry, (importlib.machinery.SourcelessFileLoader, [BC_SUFFIX]))
            spec = finder.find_spec(fullname)
            if spec is not None:
                return spec
        return None


def available():
    return reference_root() is not None


def kind():
    root = reference_root()
    if root is None:
        return None
    return "source" if root == "/root/reference" else "bytecode"


def load():
    """Put the stub and the reference on sys.path and return a namespace of its modules."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not available (neither /root/reference nor oracle/_ref)")
    try:
        importlib.import_module("gymnasium")
    except ImportError:
        if _STUB not in sys.path:
            sys.path.insert(0, _STUB)
    try:
        importlib.import_module("matplotlib.pyplot")
    except ImportError:
        if _STUB not in sys.path:
            sys.path.insert(0, _STUB)
    if root == _CANDIDATES[0]:
        if root not in sys.path:
            sys.path.insert(0, root)
    elif not any(isinstance(f, _ByteCodeFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _ByteCodeFinder(root))

    class _NS:
        pass

    ns = _NS()
    ns.root = root
    ns.envs = importlib.import_module("envs")
    ns.rewards = importlib.import_module("rewards")
    ns.policies = importlib.import_module("policies")
    ns.experiments = importlib.import_module("experiments")
    ns.episode_utils = importlib.import_module("training.episode_utils")
    ns.metrics = importlib.import_module("evaluation.metrics")
    ns.failure_taxonomy = importlib.import_module("evaluation.failure_taxonomy")
    ns.robustness_tests = importlib.import_module("evaluation.robustness_tests")
    ns.evaluator = importlib.import_module("evaluation.evaluator")
    ns.heldout_objects = importlib.import_module("evaluation.heldout_objects")
    ns.seed_variance = importlib.import_module("evaluation.seed_variance")
    ns.DexterousManipulationEnv = ns.envs.DexterousManipulationEnv
    ns.CurriculumConfig = ns.experiments.CurriculumConfig
    ns.CurriculumScheduler = ns.experiments.CurriculumScheduler
    return ns
