"""Host mirror of ``data/dataset.py:5-53`` (RangeNormalizer): pixel / frame ranges <-> [a, b].

Pure tensor plumbing on whatever device the input lives on (the kernels apply the same fp32 op
sequence internally for the hot path; this class exists for API parity: ``Tracker.range_normalizer``,
``ModelInference(range_normalizer=...)``)."""
import torch


class RangeNormalizer(torch.nn.Module):
    def __init__(self, shapes: tuple, device="cuda"):
        super().__init__()
        normalizer = torch.tensor(shapes).float().to(device) - 1
        self.register_buffer("normalizer", normalizer)

    def forward(self, x, dst=(0, 1), dims=[0, 1, 2]):
        normalized_x = x.clone()
        normalized_x[:, dims] = x[:, dims] / self.normalizer[dims]
        normalized_x[:, dims] = (dst[1] - dst[0]) * normalized_x[:, dims] + dst[0]
        return normalized_x

    def unnormalize(self, normalized_x: torch.Tensor, src=(0, 1), dims=[0, 1, 2]):
        x = normalized_x.clone()
        x[:, dims] = (normalized_x[:, dims] - src[0]) / (src[1] - src[0])
        x[:, dims] = x[:, dims] * self.normalizer[dims]
        return x
