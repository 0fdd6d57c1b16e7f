"""Mirror of source/utils/wigner_d.py:52-58 backed by the CUDA library (degrees 0..2)."""
from __future__ import annotations

import torch

from . import ops


def rotmat_to_wigner_d_matrices(max_degree: int, R: torch.Tensor):
    """R [n,3,3] -> [D_0 [n,1,1], D_1 [n,3,3], D_2 [n,5,5]][: max_degree+1] (the reference's callers drop D_0,
    source/encoder.py:249)."""
    if max_degree > 2:
        raise NotImplementedError("gta_b200: Wigner-D degrees above 2 are not implemented")
    d1, d2 = ops.wigner_d(R)
    d0 = torch.ones(R.shape[0], 1, 1, device=R.device, dtype=torch.float32)
    return [d0, d1, d2][: max_degree + 1]
