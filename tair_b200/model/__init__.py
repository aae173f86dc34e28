from .cldm import ControlLDM  # noqa: F401
from .controlnet import ControlledUnetModel, ControlNet  # noqa: F401
