"""Drop-in for the reference package of the same name (models_singlemodalinput/__init__.py:1)."""
from aide_b200.nets import UNet  # noqa: F401


def _not_built(name):
    def ctor(*a, **k):
        raise NotImplementedError(f"{name} is out of scope of the B200 hot path (SURVEY.md 2.1 #3)")
    return ctor


UNetsa = _not_built("UNetsa")
UNet2 = _not_built("UNet2"); UNet4 = _not_built("UNet4"); UNet8 = _not_built("UNet8")
UNet16 = _not_built("UNet16"); UNet32 = _not_built("UNet32"); UNet128 = _not_built("UNet128")
