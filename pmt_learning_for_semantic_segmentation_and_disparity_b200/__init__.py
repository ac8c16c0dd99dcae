"""B200-native stereo cost-volume hot path (drop-in ops for
cuevhv/PMT_learning_for_semantic_segmentation_and_disparity).

Everything here dispatches to hand-written sm_100a kernels in libpmt_ops.so through the C ABI declared in
include/pmt_ops.h.  There is no CPU, PyTorch-eager or Triton fallback: if the library is missing the ops raise.
"""
from ._lib import PmtOpsError, load as load_library  # noqa: F401
from .correlation import (CorrelationConvReLU, SpatialCorrelationSampler, SpatialCorrelationSamplerFunction,  # noqa: F401
                          correlation_conv1x1_relu, get_correlation_engine, set_correlation_engine,
                          spatial_correlation_sample)
from .psmnet import (build_concat_volume, disparityregression, matchshifted, softargmin,  # noqa: F401
                     upsample_softargmin)
from .warp import apply_disparity, photo_consistency_mse, warp_blend  # noqa: F401
from .syncbn import PairedSyncBatchNorm, PeerExchange, pair_batchnorms, sync_batchnorms_to_peer  # noqa: F401
from .compat import install_reference_shims  # noqa: F401

__version__ = "0.1.0"
