"""dlwp_cs_b200 -- B200-native cubed-sphere convolution engine (the CubeSpherePadding2D + CubeSphereConv2D hot path of
jweyn/DLWP-CS), hand-written sm_100a CUDA behind the reference's layer API.  See DESIGN.md.

    custom      CubeSpherePadding2D / CubeSphereConv2D with the reference's constructor and call signatures
    functional  the differentiable operators underneath (halo exchange, halo-fused convolution, pooling, upsample+concat)
    unet        CubeSphereUNet2 (Weyn-2020 unet2), RolloutEngine (device-resident / forced / host-streamed rollout)
    train       DataParallelTrainer (one flat NCCL all-reduce, fused Adam, CUDA-graph step, multi-step training)
    feed        DeviceDataFeed (ArrayDataGenerator on the device)
"""
from .custom import CubeSphereConv2D, CubeSpherePadding2D  # noqa: F401

__all__ = ['CubeSphereConv2D', 'CubeSpherePadding2D']
