"""Drop-in for the reference's utils package (utils/__init__.py:1-10): same names, CUDA engine behind."""
import torch
import torch.nn.functional as F

from aide_b200.losses import (CrossEntropyLoss2d, DiceLoss, CEDiceLoss, CEMDiceLoss, MulticlassDiceLoss,  # noqa: F401
                              Dice_Loss, MulticlassMSELoss, CEMDiceLossImage, Dice_fn)
from aide_b200.coteach_loss import (Coteachingloss_dropimage, Coteachingloss_dropregionce,  # noqa: F401
                                    Coteachingloss_dropimagedroppixel, Coteachingloss_weightimage)
from aide_b200.optim import PolyLR  # noqa: F401


def _hard(inputs, threshold):
    return (F.softmax(inputs, dim=1)[:, 1] >= threshold).float()


def IoU_fn(inputs, targets, threshold=0.5):                  # metrics2d.py:72-84 (reporting only)
    p, t = _hard(inputs, threshold).flatten(1), targets.float().flatten(1)
    inter = (p * t).sum(1)
    return (inter / (p.sum(1) + t.sum(1) - inter)).sum()


def TP_TN_FP_FN(inputs, targets, threshold=0.5):             # metrics2d.py:54-70 (last image only, as shipped)
    p, t = _hard(inputs, threshold)[-1].flatten(), targets[-1].float().flatten()
    return (p * t).sum(), ((1 - p) * (1 - t)).sum(), (p * (1 - t)).sum(), ((1 - p) * t).sum()


def Dice_fn_Nozero(inputs, targets, threshold=0.5):          # metrics2d.py:31-52
    p, t = _hard(inputs, threshold).flatten(1), targets.float().flatten(1)
    ps, ts = p.sum(1), t.sum(1)
    dice = torch.where(ts == 0, (ps == 0).float(), 2 * (p * t).sum(1) / (ps + ts).clamp_min(1e-30))
    count = int(((ts != 0) | (ps != 0)).sum())
    return dice.sum().item(), count


def _unsupported(name):
    def fn(*a, **k):
        raise NotImplementedError(f"{name}: multi-class reporting metric, outside the B200 hot path "
                                  "(it also crashes in the reference on numpy>=1.24, metrics2d.py:96)")
    return fn


MulticlassDice_fn = _unsupported("MulticlassDice_fn")
MulticlassIoU_fn = _unsupported("MulticlassIoU_fn")
MulticlassTP_TN_FP_FN = _unsupported("MulticlassTP_TN_FP_FN")
MulticlassAccuracy_fn = _unsupported("MulticlassAccuracy_fn")
Pixelcoreg_Focalloss = _unsupported("Pixelcoreg_Focalloss")
Pixelcoreg_Focalloss_twomodel = _unsupported("Pixelcoreg_Focalloss_twomodel")
