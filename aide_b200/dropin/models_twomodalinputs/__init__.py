"""Drop-in for the reference package of the same name (models_twomodalinputs/__init__.py:1).
Put aide_b200/dropin ahead of the reference tree on PYTHONPATH."""
from aide_b200.nets import fuseunet, fuseunetsa, fuseunetsaseparate  # noqa: F401
