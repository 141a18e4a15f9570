"""CPU restatement of single-scale SSIM as the reference's evaluation computes it through torchmetrics
(unet/scripts/output_predict.py:73,126: ``gaussian_kernel=True, kernel_size=5, sigma=0.5, data_range=1``).

torchmetrics is NOT part of the reference checkout and is not installed here: this follows its published algorithm
(``torchmetrics.functional.image.ssim._ssim_update``): Gaussian window ``exp(-(d/sigma)^2/2)`` normalised to sum 1, outer
product over the three axes, reflect padding by (k-1)/2 followed by a crop of the same border -- i.e. exactly the voxels whose
window lies inside the volume -- and C1 = (0.01 L)^2, C2 = (0.03 L)^2.  **Parity unpinned** against torchmetrics itself.
Test infrastructure only.
"""
import torch
import torch.nn.functional as F


def gaussian_window(kernel_size: int = 5, sigma: float = 0.5, dtype=torch.float64) -> torch.Tensor:
    d = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1, dtype=dtype)
    g = torch.exp(-((d / sigma) ** 2) / 2)
    g = g / g.sum()
    return (g[:, None, None] * g[None, :, None] * g[None, None, :])[None, None]


def ssim_map(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0, kernel_size: int = 5, sigma: float = 0.5):
    """x, y: [N, 1, D, H, W]; returns the SSIM map over the valid region [N, 1, D-4, H-4, W-4]."""
    w = gaussian_window(kernel_size, sigma, x.dtype)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mx, my = F.conv3d(x, w), F.conv3d(y, w)
    sxx, syy, sxy = F.conv3d(x * x, w), F.conv3d(y * y, w), F.conv3d(x * y, w)
    vx, vy, cxy = sxx - mx * mx, syy - my * my, sxy - mx * my
    return ((2 * mx * my + c1) * (2 * cxy + c2)) / ((mx * mx + my * my + c1) * (vx + vy + c2))


def ssim_loss(x: torch.Tensor, y: torch.Tensor, **kw) -> torch.Tensor:
    return 1.0 - ssim_map(x, y, **kw).mean()


MS_SSIM_BETAS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def ms_ssim(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0, kernel_size: int = 5, sigma: float = 0.5) -> torch.Tensor:
    """torchmetrics ``multiscale_structural_similarity_index_measure`` (normalize="relu", default betas), restated:
    per image prod_i relu(cs_i)^beta_i with the last factor replaced by relu(ssim), avg_pool3d(2) between scales."""
    w = gaussian_window(kernel_size, sigma, x.dtype)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    terms = []
    for i, beta in enumerate(MS_SSIM_BETAS):
        mx, my = F.conv3d(x, w), F.conv3d(y, w)
        sxx, syy, sxy = F.conv3d(x * x, w), F.conv3d(y * y, w), F.conv3d(x * y, w)
        vx, vy, cxy = sxx - mx * mx, syy - my * my, sxy - mx * my
        cs = (2 * cxy + c2) / (vx + vy + c2)
        sim = ((2 * mx * my + c1) / (mx * mx + my * my + c1)) * cs
        last = i == len(MS_SSIM_BETAS) - 1
        val = (sim if last else cs).flatten(1).mean(1)
        terms.append(torch.relu(val) ** beta)
        if not last:
            x, y = F.avg_pool3d(x, 2), F.avg_pool3d(y, 2)
    return torch.stack(terms).prod(0).mean()
