"""scattennet_b200 - B200-native (sm_100a) SCAttenNet spatial-coordinate
attention encoder behind the reference's module interfaces.

The nn.Module classes mirror ``model/{attention,encoder,fusion,keypoint_module,
layers,residual,utils}.py`` of tinh2044/SCAttenNet (constructor arguments,
forward signatures, state-dict layout); the arithmetic runs in hand-written
CUDA kernels (``csrc/``) reached through the C ABI of ``libscatt.so``
(``include/scatt.h``).  There is no CPU fallback.
"""

from .alignment_module import AlignmentModule  # noqa: F401
from .attention import BaseAttention, CrossAttention, SelfAttention, SelfCausalAttention  # noqa: F401
from .decode import ctc_decode  # noqa: F401
from .config import PHOENIX_2014, PHOENIX_2014T, model_config  # noqa: F401
from .encoder import Encoder, EncoderLayer  # noqa: F401
from .functional import PRECISIONS, get_precision, set_default_precision  # noqa: F401
from .fusion import CoordinatesFusion, InvertedResidual  # noqa: F401
from .keypoint_module import (  # noqa: F401
    CoordinateAttention,
    CoordinatesMerge,
    KeypointModule,
    SeparativeCoordinateAttention,
)
from .layers import CoordinateMapping, FeedForward, LearningPositionEmbedding  # noqa: F401
from .model import LinearHeads, MSCAEncoder  # noqa: F401
from .residual import ResidualBlock, ResidualNetwork  # noqa: F401

__version__ = "0.1.0"
