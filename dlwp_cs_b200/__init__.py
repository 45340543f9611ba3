"""dlwp_cs_b200 -- B200-native cubed-sphere convolution engine (the CubeSpherePadding2D + CubeSphereConv2D hot path of
jweyn/DLWP-CS), hand-written sm_100a CUDA behind the reference's layer API.  See DESIGN.md."""
from .custom import CubeSphereConv2D, CubeSpherePadding2D  # noqa: F401

__all__ = ['CubeSphereConv2D', 'CubeSpherePadding2D']
