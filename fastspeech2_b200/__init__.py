"""fastspeech2_b200 -- B200-native (sm_100a) FastSpeech2 mel-synthesis forward path behind the
reference's `FeedForwardTransformer` API.  See DESIGN.md / INTEGRATION.md."""
from .fastspeech import FeedForwardTransformer  # noqa: F401
from .length_regulator import LengthRegulator  # noqa: F401
from .weights import ModelDims, synthetic_state_dict  # noqa: F401

__all__ = ["FeedForwardTransformer", "LengthRegulator", "ModelDims", "synthetic_state_dict"]
