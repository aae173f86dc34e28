from .models import TESTR, MSDeformAttn  # noqa: F401
from .structures import Instances  # noqa: F401
from .transformer_detector import TransformerDetector, default_cfg  # noqa: F401
