"""osufusion_b200 — B200-native (sm_100a) implementation of OsuFusion's denoiser hot path.

Public API mirrors the reference: `osufusion_b200.models.diffusion.OsuFusion`, `osufusion_b200.models.rectified_flow.OsuFusion`,
`osufusion_b200.modules.UNet`.  All compute runs in libosufusion_sm100.so (hand-written CUDA behind include/osufusion_b200.h).
"""
__all__ = ["modules", "models", "engine"]
