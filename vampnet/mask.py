"""vampnet.mask (reference vampnet/mask.py) -> vampnet_b200.mask."""
from vampnet_b200.mask import *  # noqa: F401,F403
from vampnet_b200.mask import _gamma, _invgamma  # noqa: F401  (train.py uses pmask._gamma)
