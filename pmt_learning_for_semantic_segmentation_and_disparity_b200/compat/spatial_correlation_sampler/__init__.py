"""`import spatial_correlation_sampler` resolves here when .../compat is on PYTHONPATH."""
from pmt_learning_for_semantic_segmentation_and_disparity_b200.correlation import (  # noqa: F401
    SpatialCorrelationSampler, SpatialCorrelationSamplerFunction, spatial_correlation_sample)

__version__ = "b200-0.1.0"
