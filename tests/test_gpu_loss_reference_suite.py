"""The reference's whole unit-test file, case for case, against the drop-in evaluator on CUDA tensors.

/root/reference/test/loss/test_RegressionLossEvaluator.py:7-193 holds 24 cases on the four static helpers
(get_squared_diff_mean_vector 9-35, get_mask_by_threes 37-87, get_mean_norm_error 89-159, get_com_acc_error
161-193).  Every case is restated here with the reference's literal tensors, shapes ((2,4,3), (1,2,3), (2,2,3),
(1,1,6) …), thresholds and assertion kinds (torch.equal / allclose / isclose / assertRaises(ValueError)); the only
change is ``.cuda()`` on the inputs (the drop-in has no CPU path) and ``.cpu()`` on results before comparing.
Test ids carry the reference's method names so the two files can be read side by side.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    from inferbiomechanics_b200.loss.RegressionLossEvaluator import RegressionLossEvaluator
    return RegressionLossEvaluator


def g(x):
    return (x if isinstance(x, torch.Tensor) else torch.tensor(x)).cuda()


# ---- get_squared_diff_mean_vector (ref 9-35) -----------------------------------------------------------------
def test_get_squared_diff_mean_vector_with_valid_tensors(R):
    o = torch.tensor(range(24)).reshape((2, 4, 3)) * 1.
    got = R.get_squared_diff_mean_vector(g(o), g(o.clone()))
    assert torch.equal(got.cpu(), torch.tensor([0.0, 0.0, 0.0]))


def test_get_squared_diff_mean_vector_with_nonzero_loss(R):
    o = torch.tensor(range(24)).reshape((2, 4, 3)) * 1.
    got = R.get_squared_diff_mean_vector(g(o), g(o + 1.))
    assert torch.allclose(got.cpu(), torch.tensor([1., 1., 1.]))


def test_get_squared_diff_mean_vector_with_mismatched_tensor_shapes(R):
    with pytest.raises(ValueError):
        R.get_squared_diff_mean_vector(g([[[1.0, 2.0], [3.0, 4.0]]]), g([[[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]]))


def test_get_squared_diff_mean_vector_with_empty_tensors(R):
    with pytest.raises(ValueError):
        R.get_squared_diff_mean_vector(g(torch.tensor([])), g(torch.tensor([])))


# ---- get_mask_by_threes (ref 37-87) ----------------------------------------------------------------------------
def test_mask_by_threes_with_valid_input(R):
    x = [[[1.0, 0.0, 0.0], [0.0, 2.0, 0.0]], [[0., 0., 0.], [3., 0., 4.]]]
    want = torch.tensor([[[1.0, 1.0, 1.0], [1.0, 1.0, 1.0]], [[0., 0., 0.], [1., 1., 1.]]])
    assert torch.equal(R.get_mask_by_threes(g(x)).cpu(), want)


def test_mask_by_threes_with_theshold(R):
    x = [[[1.0, 0.0, 0.0], [0.0, 2.0, 0.0]]]
    want = torch.tensor([[[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]]])
    assert torch.equal(R.get_mask_by_threes(g(x), threshold=1.5).cpu(), want)


def test_mask_by_threes_with_invalid_dimensions(R):
    with pytest.raises(ValueError):
        R.get_mask_by_threes(g([[1.0, 0.0, 0.0]]))


def test_mask_by_threes_with_empty_tensor(R):
    with pytest.raises(ValueError):
        R.get_mask_by_threes(g(torch.empty(0)))


def test_mask_by_threes_with_invalid_last_dimension(R):
    with pytest.raises(ValueError):
        R.get_mask_by_threes(g([[[1.0, 0.0], [0.0, 2.0]]]))


def test_mask_by_threes_with_zeros(R):
    x = [[[0.0, 0.0, 0.0], [0.0, 0.0, 0.0]]]
    assert torch.equal(R.get_mask_by_threes(g(x)).cpu(), torch.tensor(x))


def test_mask_by_threes_with_one_non_zero(R):
    x = [[[0.0, 0.0, 1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0, 0.0, 0.0]]]
    want = torch.tensor([[[1.0, 1.0, 1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0, 1.0, 1.0]]])
    assert torch.equal(R.get_mask_by_threes(g(x)).cpu(), want)


# ---- get_mean_norm_error (ref 89-159) --------------------------------------------------------------------------
def test_get_mean_norm_error_shape_mismatch(R):
    with pytest.raises(ValueError):
        R.get_mean_norm_error(g(torch.rand((3, 2, 6))), g(torch.rand((3, 2, 9))))


def test_get_mean_norm_error_tensor_not_3d(R):
    with pytest.raises(ValueError):
        R.get_mean_norm_error(g(torch.rand(2, 6)), g(torch.rand(2, 6)))


def test_get_mean_norm_error_empty_tensor(R):
    with pytest.raises(ValueError):
        R.get_mean_norm_error(g(torch.rand(0, 6)), g(torch.rand(0, 6)))


def test_get_mean_norm_error_final_dimension_not_divisible_by_three(R):
    with pytest.raises(ValueError):
        R.get_mean_norm_error(g(torch.rand(3, 2, 7)), g(torch.rand(3, 2, 7)))


LAB_223 = [[[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]]


def test_get_mean_norm_error_zero(R):
    o = [[[3., -2.0, 4.0], [4.0, 5.0, 6.0]], [[3., -2.0, 4.0], [4.0, 5.0, 6.0]]]      # differs in frame 0 only
    got = R.get_mean_norm_error(g(o), g(LAB_223))
    assert torch.isclose(got.cpu(), torch.tensor([0.]))


def test_get_mean_norm_error_non_zero(R):
    o = [[[3., -2.0, 4.0], [4.0, 5.0, 6.0]], [[3., -2.0, 4.0], [4.0, 5.0, 7.0]]]      # last frame: one vector off by 1
    got = R.get_mean_norm_error(g(o), g(LAB_223))
    assert torch.isclose(got.cpu(), torch.tensor([0.5]))


def test_get_mean_norm_error_zero_vec_size_6(R):
    v = [[[1.0, 2.0, 3.0, 4.0, 5.0, 6.0]]]
    got = R.get_mean_norm_error(g(v), g(v), vec_size=6)
    assert torch.isclose(got.cpu(), torch.tensor([0.0]))


def test_get_mean_norm_error_non_zero_vec_size_6(R):
    v = [[[1.0, 2.0, 3.0, 4.0, 5.0, 6.0]]]
    got = R.get_mean_norm_error(g(v), g([[[0.0] * 6]]), vec_size=6)
    assert torch.isclose(got.cpu(), torch.norm(torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])))


# ---- get_com_acc_error (ref 161-193) ---------------------------------------------------------------------------
def test_shape_mismatch(R):
    with pytest.raises(ValueError):
        R.get_com_acc_error(g(torch.rand(3, 2, 6)), g(torch.rand(4, 2, 6)))


def test_not_3d_tensor(R):
    with pytest.raises(ValueError):
        R.get_com_acc_error(g(torch.rand(2, 6)), g(torch.rand(2, 6)))


def test_empty_tensor(R):
    with pytest.raises(ValueError):
        R.get_com_acc_error(g(torch.empty(0, 0)), g(torch.rand(3, 6)))


def test_final_dimension_not_six(R):
    with pytest.raises(ValueError):
        R.get_com_acc_error(g(torch.rand(3, 2, 5)), g(torch.rand(3, 2, 5)))


def test_output_zero(R):
    o = [[[1.0, 2.0, 3.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0, 2.0, 3.0]]]
    l = [[[0.0, 0.0, 0.0, 1.0, 2.0, 3.0], [1.0, 2.0, 3.0, 0.0, 0.0, 0.0]]]            # left/right swapped
    got = R.get_com_acc_error(g(o), g(l))
    assert torch.isclose(got.cpu(), torch.tensor([0.0]))


# ---- beyond the reference's cases: the helpers against the oracle restatement at odd shapes, and autograd --------
@pytest.mark.parametrize("B,F,C,v", [(5, 7, 9, 3), (3, 1, 12, 4), (130, 50, 30, 6), (2, 3, 201, 67), (4097, 10, 6, 3)])
def test_helpers_general_shapes_match_oracle(R, B, F, C, v):
    from oracle import loss as oloss
    gen = torch.Generator().manual_seed(B * 1000 + C)
    o = torch.randn(B, F, C, generator=gen) * 3
    l = torch.randn(B, F, C, generator=gen) * 3
    got = R.get_squared_diff_mean_vector(g(o), g(l)).cpu()
    assert torch.allclose(got, oloss.squared_diff_mean_vector(o, l), rtol=2e-6, atol=1e-7)      # fp32 kernel bar
    got = R.get_mean_norm_error(g(o), g(l), vec_size=v).cpu()
    assert torch.allclose(got, oloss.mean_norm_error(o, l, vec_size=v), rtol=2e-6, atol=1e-7)
    if C % 3 == 0:
        for thr in (0.0, 2.5, 10.0):
            assert torch.equal(R.get_mask_by_threes(g(l), threshold=thr).cpu(), oloss.mask_by_threes(l, threshold=thr))
    # non-contiguous (B, F) strides: a slice of a wider buffer
    wide = torch.randn(B, F, C + 5, generator=gen)
    ov = g(wide)[:, :, 2:2 + C]
    got = R.get_squared_diff_mean_vector(ov, g(l)).cpu()
    assert torch.allclose(got, oloss.squared_diff_mean_vector(wide[:, :, 2:2 + C], l), rtol=2e-6, atol=1e-7)


def test_squared_diff_mean_vector_autograd(R):
    gen = torch.Generator().manual_seed(5)
    o = torch.randn(6, 4, 9, generator=gen)
    l = torch.randn(6, 4, 9, generator=gen)
    up = torch.randn(9, generator=gen)
    oc = o.clone().requires_grad_(True)
    (((oc - l) ** 2).mean(dim=(0, 1)) * up).sum().backward()
    og = g(o).requires_grad_(True)
    (R.get_squared_diff_mean_vector(og, g(l)) * g(up)).sum().backward()
    assert torch.allclose(og.grad.cpu(), oc.grad, rtol=2e-6, atol=1e-7)
