"""Image losses of the reference's optimisation loop (reference examples/test_opt_tf.py:70-72):

    dssim_loss = 1.0 - ssim2d(res, gt, data_range=1.0, size_average=True, nonnegative_ssim=True)     # pytorch_msssim.ssim
    mse_loss   = F.mse_loss(res, gt)
    loss       = torch.nan_to_num(dssim_loss) + mse_loss

`pytorch_msssim` is not installed here, so the SSIM it computes (Wang et al. 2004 as implemented there: separable 11-tap
Gaussian window, sigma 1.5, 'valid' filtering, K = (0.01, 0.03), per-channel mean, optional ReLU) is restated with plain torch
ops.  The MSE half can come fused out of the ray-march itself (`Raycaster.mse_loss`: loss in the forward epilogue, gradient
formed inside the backward kernel); the DSSIM half works on the rendered image -- a few library convolutions over an image, not
part of the march -- and reaches the march through the ordinary autograd backward of `Raycaster.forward`.
"""
import torch
import torch.nn.functional as F

__all__ = ["ssim", "dssim_loss", "mse_dssim_loss"]


def _gauss_window(size, sigma, dtype, device):
    x = torch.arange(size, dtype=dtype, device=device) - size // 2
    g = torch.exp(-(x ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filter(x, win):
    """Separable 'valid' Gaussian blur of (N, C, H, W), one window per channel (an axis shorter than the window is left alone)."""
    C = x.shape[1]
    k = win.numel()
    if x.shape[2] >= k:
        x = F.conv2d(x, win.view(1, 1, k, 1).expand(C, 1, k, 1), groups=C)
    if x.shape[3] >= k:
        x = F.conv2d(x, win.view(1, 1, 1, k).expand(C, 1, 1, k), groups=C)
    return x


def ssim(X, Y, data_range=1.0, size_average=True, win_size=11, win_sigma=1.5, K=(0.01, 0.03), nonnegative_ssim=False):
    """Structural similarity of two (N, C, H, W) image batches, as pytorch_msssim.ssim computes it."""
    if X.shape != Y.shape or X.ndim != 4:
        raise ValueError(f"ssim expects two (N, C, H, W) tensors of equal shape, got {tuple(X.shape)} and {tuple(Y.shape)}")
    Y = Y.to(X.dtype)
    win = _gauss_window(win_size, win_sigma, X.dtype, X.device)
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _filter(X, win), _filter(Y, win)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _filter(X * X, win) - mu1_sq
    s2 = _filter(Y * Y, win) - mu2_sq
    s12 = _filter(X * Y, win) - mu12
    cs = (2 * s12 + C2) / (s1 + s2 + C2)
    ssim_map = ((2 * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs
    per_channel = ssim_map.flatten(2).mean(-1)
    if nonnegative_ssim:
        per_channel = torch.relu(per_channel)
    return per_channel.mean() if size_average else per_channel.mean(1)


def dssim_loss(pred, target):
    """1 - SSIM with the reference's arguments (examples/test_opt_tf.py:70)."""
    return 1.0 - ssim(pred, target, data_range=1.0, size_average=True, nonnegative_ssim=True)


def mse_dssim_loss(pred, target):
    """The reference's training loss (examples/test_opt_tf.py:70-72): nan_to_num(DSSIM) + MSE."""
    return torch.nan_to_num(dssim_loss(pred, target)) + F.mse_loss(pred, target)
