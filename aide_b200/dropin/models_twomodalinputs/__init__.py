"""Drop-in for the reference package of the same name (models_twomodalinputs/__init__.py:1).
Put aide_b200/dropin ahead of the reference tree on PYTHONPATH."""
from aide_b200.nets import fuseunet  # noqa: F401


def _not_built(name):
    def ctor(*a, **k):
        raise NotImplementedError(f"{name} (attention variant) is out of scope of the B200 hot path (SURVEY.md 2.1 #1)")
    return ctor


fuseunetsa = _not_built("fuseunetsa")
fuseunetsaseparate = _not_built("fuseunetsaseparate")
