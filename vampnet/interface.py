"""vampnet.interface (reference vampnet/interface.py) -> vampnet_b200.interface."""
from vampnet_b200.interface import *  # noqa: F401,F403  (also brings the mask helpers, like the reference's `from .mask import *`)
from vampnet_b200.interface import Interface, signal_concat, _load_model  # noqa: F401
from vampnet_b200.audio import AudioSignal  # noqa: F401  (the reference module has `from audiotools import AudioSignal`)
from vampnet_b200.codec import DAC  # noqa: F401  (reference: `from lac.model.lac import LAC as DAC`, interface.py:16)
