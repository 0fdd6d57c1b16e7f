"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, via oracle/ref_harness.py) in the build container.

    python tests/golden/gen_golden.py

The reference has no tests, fixtures or known-answer vectors of its own, so these
files are what pins parity: inputs + the reference's fp32 outputs on CPU.  Re-running this script
reproduces the files bit-for-bit (the reference is deterministic on CPU).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gta_b200.synth import (CFG1_A, CFG1_B, CLEVR, CLEVR_EUCLID, CLEVR_T2, MSN_SO3, MSN_SO3_EUCLID, MSN_T2,  # noqa: E402
                            GtaConfig, make_inputs)
from oracle import ref_harness as rh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, base cfg, Nq, Nk, tq/view, tk/view, cross, B, trans_coeff, seed, v_transform
    ("cfg1b_self", CFG1_B, 2, 2, 16, 16, False, 2, 0.01, 11, True),
    ("cfg1a_self_tc1", CFG1_A, 2, 2, 16, 16, False, 1, 1.0, 12, True),
    ("msn_cross", MSN_SO3, 3, 2, 8, 16, True, 1, 0.01, 13, True),
    ("msn_self_tc1", MSN_SO3, 5, 5, 4, 4, False, 1, 1.0, 14, True),
    ("clevr_self_ragged", CLEVR, 2, 2, 21, 21, False, 1, 0.01, 15, True),
    ("clevr_cross_novt", CLEVR, 3, 2, 7, 12, True, 1, 0.01, 16, False),
]
# ablation configs (t2 block, euclid_sim, blocks that are not multiples of 8): same tuple layout
ABLATION_CASES = [
    ("clevr_t2_self", CLEVR_T2, 2, 2, 21, 21, False, 1, 0.01, 31, True),
    ("msn_t2_cross_tc1", MSN_T2, 3, 2, 8, 16, True, 1, 1.0, 32, True),
    ("clevr_euclid_self", CLEVR_EUCLID, 2, 2, 21, 21, False, 1, 0.01, 33, True),
    ("msn_so3_euclid_cross", MSN_SO3_EUCLID, 3, 2, 8, 16, True, 1, 0.3, 34, True),
    ("clevr_euclid_cross_novt", CLEVR_EUCLID, 3, 2, 7, 12, True, 1, 0.01, 35, False),
]


def gimbal_extrinsics():
    """Views whose inv(E) rotation hits the two gimbal branches (R22 = +1 / -1) plus identity and a
    generic one (wigner_d.py:44-48)."""
    def rz(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    rx_pi = np.diag([1.0, -1.0, -1.0])
    gen = torch.Generator().manual_seed(5)
    A = torch.randn(3, 3, generator=gen, dtype=torch.float64)
    Q, _ = torch.linalg.qr(A)
    if torch.linalg.det(Q) < 0:
        Q[:, 0] = -Q[:, 0]
    Rs = [np.eye(3), rz(0.7), rz(-2.1) @ rx_pi, Q.numpy(), rz(3.0)]
    E = np.zeros((1, len(Rs), 4, 4), np.float32)
    for i, R in enumerate(Rs):
        E[0, i, :3, :3] = R.T          # E = inverse pose, so inv(E) has rotation R
        E[0, i, :3, 3] = [0.3 * i, -0.2, 0.5]
        E[0, i, 3, 3] = 1
    return torch.from_numpy(E)


def main():
    for name, base, nq, nk, tq, tk, cross, B, tc, seed, vt in CASES + ABLATION_CASES:
        cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk, v_transform=vt)
        inp = make_inputs(cfg, B, tq, tk, cross=cross, seed=seed)
        out, ex = rh.ref_gta_attention(cfg, inp, trans_coeff=tc)
        d = {k: v.contiguous().numpy() for k, v in inp.items()}
        d["out"] = out.contiguous().numpy()
        d["trans_coeff"] = np.float32(tc)
        d["cross"] = np.int32(cross)
        for key in ("se3rep_q", "se3rep_k", "inv_se3rep_q", "so2rep_q", "so2rep_k", "t2rep_q", "t2rep_k", "inv_t2rep_q"):
            if key in ex:
                d["ref_" + key] = ex[key].contiguous().numpy()
        for key in ("so3rep_q", "so3rep_k"):
            if key in ex:
                d["ref_" + key + "_d1"] = ex[key][0].contiguous().numpy()
                d["ref_" + key + "_d2"] = ex[key][1].contiguous().numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, d["out"].shape, float(np.abs(d["out"]).mean()))

    # rep-only golden incl. gimbal cases
    m = rh.load()
    E = gimbal_extrinsics()
    R = torch.linalg.inv(E)[..., :3, :3].flatten(0, 1)
    Ds = m.wigner_d.rotmat_to_wigner_d_matrices(2, R)
    coord = torch.rand(1, 40, 2, generator=torch.Generator().manual_seed(6))
    so2 = m.gta.make_SO2mats(coord, 6, [1, 1]).flatten(-4, -3)
    so2b = m.gta.make_SO2mats(coord, 3, [2, 0.5], shared_freqs=True).flatten(-4, -3)
    np.savez_compressed(os.path.join(HERE, "reps_gimbal.npz"), extr=E.numpy(), inv=torch.linalg.inv(E).numpy(),
                        d0=Ds[0].numpy(), d1=Ds[1].numpy(), d2=Ds[2].numpy(), coord=coord.numpy(),
                        so2_n6=so2.numpy(), so2_n3_shared_f2_05=so2b.numpy(),
                        coord2d_5x7=m.gta.make_2dcoord(5, 7))
    print("reps_gimbal", Ds[1].shape, Ds[2].shape, so2.shape)


if __name__ == "__main__":
    main()
