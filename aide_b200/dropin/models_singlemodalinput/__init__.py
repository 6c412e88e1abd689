"""Drop-in for the reference package of the same name (models_singlemodalinput/__init__.py:1):
``from models_singlemodalinput import UNet, UNetsa`` (train_files/trainkidney_proposed_mask1.py:23) resolves here when
aide_b200/dropin precedes the reference tree on PYTHONPATH."""
from aide_b200.nets import UNet, UNetsa, UNet2, UNet4, UNet8, UNet16, UNet32, UNet128  # noqa: F401
