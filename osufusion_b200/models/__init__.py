from .diffusion import OsuFusion as DiffusionOsuFusion  # noqa: F401
from .rectified_flow import OsuFusion as RectifiedFlowOsuFusion  # noqa: F401
