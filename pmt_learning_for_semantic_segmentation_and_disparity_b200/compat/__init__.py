"""Make the reference's imports resolve to the B200 ops.

Either put ``<repo>/pmt_learning_for_semantic_segmentation_and_disparity_b200/compat`` on PYTHONPATH (it contains a
``spatial_correlation_sampler`` package with the upstream names), or call :func:`install_reference_shims` before
importing the reference's ``models`` package.
"""
from __future__ import annotations

import sys
import types


def install_reference_shims(patch_reference_modules: bool = True) -> None:
    """Register `spatial_correlation_sampler` in sys.modules and, if the reference's modules are already
    imported, rebind their hot-path symbols (apply_disparity, disparityregression, matchshifted)."""
    from .. import correlation, psmnet, warp

    m = types.ModuleType("spatial_correlation_sampler")
    m.SpatialCorrelationSampler = correlation.SpatialCorrelationSampler
    m.SpatialCorrelationSamplerFunction = correlation.SpatialCorrelationSamplerFunction
    m.spatial_correlation_sample = correlation.spatial_correlation_sample
    m.__version__ = "b200-0.1.0"
    sys.modules["spatial_correlation_sampler"] = m
    if not patch_reference_modules:
        return
    td = sys.modules.get("models.torch_dsnet")
    if td is not None:
        td.apply_disparity = warp.apply_disparity
    for name in ("models.dsnet_t2_warp", "models.dsnet_t2"):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "apply_disparity"):
            mod.apply_disparity = warp.apply_disparity
    sm = sys.modules.get("models_psmnet.submodule")
    if sm is not None:
        sm.disparityregression = psmnet.disparityregression
        sm.matchshifted = psmnet.matchshifted
    sh = sys.modules.get("models_psmnet.stackhourglass")
    if sh is not None and hasattr(sh, "disparityregression"):
        sh.disparityregression = psmnet.disparityregression
