"""-m gpu: empty inputs through every op of the hot path (the reference's ATen ops accept empty batches)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def pmt():
    import pmt_learning_for_semantic_segmentation_and_disparity_b200 as m
    m.load_library()
    return m


def test_empty_batch_everywhere(pmt):
    z = lambda *s: torch.zeros(*s, device=DEV)
    a, b = z(0, 8, 4, 16).requires_grad_(True), z(0, 8, 4, 16).requires_grad_(True)
    out = pmt.SpatialCorrelationSampler(kernel_size=1, patch_size=(1, 5), stride=1, padding=0, dilation_patch=1)(a, b)
    assert out.shape == (0, 1, 5, 4, 16)
    out.sum().backward()
    assert a.grad.shape == a.shape and b.grad.shape == b.shape
    assert pmt.build_concat_volume(z(0, 4, 6, 32), z(0, 4, 6, 32), 7).shape == (0, 8, 7, 6, 32)
    assert pmt.softargmin(z(0, 24, 6, 32)).shape == (0, 6, 32)
    assert pmt.disparityregression(24)(z(0, 24, 6, 32)).shape == (0, 6, 32)
    assert pmt.upsample_softargmin(z(0, 1, 6, 3, 8), 24, (12, 32)).shape == (0, 12, 32)
    assert pmt.apply_disparity(z(0, 3, 6, 32), z(0, 1, 6, 32)).shape == (0, 3, 6, 32)


def test_zero_rows_and_single_pixel(pmt):
    a, b = torch.randn(1, 4, 1, 4, device=DEV), torch.randn(1, 4, 1, 4, device=DEV)
    out = pmt.SpatialCorrelationSampler(kernel_size=1, patch_size=(1, 1), stride=1, padding=0, dilation_patch=1)(a, b)
    assert torch.allclose(out[0, 0, 0], (a * b).sum(1)[0], atol=1e-5)
    cost = torch.randn(1, 1, 1, 1, device=DEV)
    assert float(pmt.softargmin(cost)) == 0.0      # a single plane: probability 1 at disparity 0


def test_even_patch_with_dilation_is_refused(pmt):
    """Upstream's CPU and CUDA builds centre an even, dilated patch differently; no reference call site uses it, so the
    op refuses it (Python surface and C ABI) instead of guessing."""
    import ctypes

    a = torch.zeros(1, 2, 4, 8, device="cuda:0")
    with pytest.raises(NotImplementedError):
        pmt.spatial_correlation_sample(a, a, patch_size=(1, 4), dilation_patch=2)
    lib = pmt.load_library()
    out = torch.empty(1, 1, 4, 4, 8, device="cuda:0")
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.pmt_corr_fwd_f32(vp(a), vp(a), vp(out), 1, 2, 4, 8, 1, 4, 1, 2, None) == 3   # PMT_ERR_UNSUPPORTED
    assert b"even patch" in lib.pmt_last_error()
