"""Evaluation metric of the reference's ``src/loss.py``: ``EPE`` (:12-21), the end-point error every loss class there
reports next to its training loss and the number PIV accuracy is quoted in (AEE).  Works on the tensors the drop-in
models return, on whatever device they live; the training losses themselves (L1 / L2 / multi-scale) are out of scope."""
import torch

__all__ = ["EPE"]


def EPE(input_flow: torch.Tensor, target_flow: torch.Tensor, mean: bool = True) -> torch.Tensor:
    """Per-pixel Euclidean distance between two ``[B, 2, H, W]`` flows, averaged over all pixels (``mean``) or summed and
    divided by the batch size."""
    diff = target_flow - input_flow
    epe_map = torch.sqrt((diff * diff).sum(dim=1))
    return epe_map.mean() if mean else epe_map.sum() / epe_map.size(0)
