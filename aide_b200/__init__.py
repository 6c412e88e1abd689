"""aide_b200 -- B200-native (sm_100a) engine for the data-parallel hot path of lich0031/AIDE.

Public surface (mirrors the reference's Python modules so train_files/*.py can import it as a drop-in;
see INTEGRATION.md):

    from aide_b200 import fuseunet, UNet                      # models_twomodalinputs / models_singlemodalinput
    from aide_b200 import CEMDiceLossImage, DiceLoss, ...     # utils.loss2d
    from aide_b200 import Dice_fn                             # utils.metrics2d
    from aide_b200 import pseudo_label, coteach_step          # fused form of the inline AIDE step

Importing this package loads aide_b200/libaide_b200.so and raises ImportError if it has not been
built: there is no CPU or PyTorch fallback for the compute path.
"""
from ._lib import AideError, FMT_BF16, FMT_F16X2, FMT_F32, FMT_TF32X2, LIB_PATH, lib  # noqa: F401
from .engine import MODES, default_mode  # noqa: F401
from .nets import (UNet, UNet4, UNet8, UNet16, UNet32, UNet128, UNetsa, fuseunet, fuseunetsa,  # noqa: F401
                   fuseunetsaseparate)
from .losses import (CEDiceLoss, CEMDiceLoss, CEMDiceLossImage, CrossEntropyLoss2d, Dice_Loss, DiceLoss,  # noqa: F401
                     Dice_fn, MulticlassDiceLoss, MulticlassMSELoss, coteach_step, predict_mask, pseudo_label)
from .coteach_loss import (Coteachingloss_dropimage, Coteachingloss_dropimagedroppixel,  # noqa: F401
                           Coteachingloss_dropregionce, Coteachingloss_weightimage)
from .optim import FlatAdamAMSGrad, PolyLR  # noqa: F401
from .augment import augmented_views, forward_aug, reverse_aug_tensor, reverseaug  # noqa: F401

__version__ = "0.1.0"
