"""dlwp_cs_b200 -- B200-native cubed-sphere convolution engine (the CubeSpherePadding2D + CubeSphereConv2D hot path of
jweyn/DLWP-CS), hand-written sm_100a CUDA behind the reference's layer API.  See DESIGN.md.

    custom      CubeSpherePadding2D / CubeSphereConv2D with the reference's constructor and call signatures
    functional  the differentiable operators underneath (halo exchange, halo-fused convolution, pooling, upsample+concat)
    unet        CubeSphereCNN (basic / unet / unet2 / unet3 / unet4), RolloutEngine (device-resident / forced / host-streamed)
    models      CubeSphereForecaster: predict / predict_timeseries with the reference's arguments and array layout
    h5weights   pure-Python HDF5 reader for the reference's saved Keras models
    train       DataParallelTrainer (flat gradient buffer, bucketed all-reduce overlapped with backward, fused Adam, graph)
    feed        DeviceDataFeed (ArrayDataGenerator on the device)
"""
from .custom import CubeSphereConv2D, CubeSpherePadding2D  # noqa: F401

__all__ = ['CubeSphereConv2D', 'CubeSpherePadding2D']
