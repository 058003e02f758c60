"""CPU tests (SIMT emulator) of the crop / resample kernels and their host code against the numpy restatement of the
reference's collect_pc / resampling (oracle/grasp_crop_np.py)."""
import numpy as np
import pytest
import torch

import emu_util as E
from oracle import grasp_crop_np as OC
from oracle import weights as W
from pointnetgpd_b200 import prep


def _scene(seed, P=3000, G=7):
    pc = (W.uniform(seed, (P, 3), -0.12, 0.12)).astype(np.float32)
    centers = W.uniform(seed + 1, (G, 3), -0.05, 0.05)
    axes = W.normal(seed + 2, (G, 3))
    axes[0] = [0, 0, 1.0]                                  # degenerate case of dataset.py:29-30
    width = W.uniform(seed + 3, (G,), 0.05, 0.085)
    angle = W.uniform(seed + 4, (G,), -1.5, 1.5)
    grasps = np.concatenate([centers, axes, width[:, None], angle[:, None], np.zeros((G, 4))], axis=1)
    th = 0.3
    T = np.eye(4)
    T[:3, :3] = [[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]]
    T[:3, 3] = [0.01, -0.02, 0.005]
    return pc, grasps, T


def test_frames_match_reference_formulas():
    pc, grasps, T = _scene(1)
    fr = prep.grasp_frames(grasps, T)
    for g in range(len(grasps)):
        c, M, w = OC.grasp_frame(grasps[g], T)
        assert np.allclose(fr[g, 0:3], c, atol=1e-14)
        assert np.allclose(fr[g, 3:12].reshape(3, 3), M, atol=1e-13)
        assert np.allclose(fr[g, 12:15], [w / 4, w / 2, w / 4])


def test_crop_is_index_exact():
    pc, grasps, T = _scene(2)
    lib = E.emu_lib()
    offsets, pts, idx = prep.crop(torch.tensor(pc), prep.grasp_frames(grasps, T), lib=lib)
    assert int(offsets[-1]) > 50
    for g in range(len(grasps)):
        ref_idx, ref_pts = OC.crop(pc, grasps[g], T)
        a, b = int(offsets[g]), int(offsets[g + 1])
        assert np.array_equal(idx[a:b].numpy(), ref_idx), g          # integer/index work: bit exact
        assert np.allclose(pts[a:b].numpy(), ref_pts.astype(np.float32), atol=1e-7)


@pytest.mark.parametrize("N", [64, 500])
def test_resample_properties(N):
    lib = E.emu_lib()
    sizes = [0, 5, 63, 64, 65, 700, 1500]
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
    pts = torch.tensor(W.normal(5, (sum(sizes), 3)).astype(np.float32))
    x, oi = prep.resample(pts, offsets, N, repeat=3, seed=42, return_index=True, lib=lib)
    x2, oi2 = prep.resample(pts, offsets, N, repeat=3, seed=42, return_index=True, lib=lib)
    assert torch.equal(x, x2) and torch.equal(oi, oi2)               # deterministic in the seed
    x3, oi3 = prep.resample(pts, offsets, N, repeat=3, seed=43, return_index=True, lib=lib)
    assert not torch.equal(oi[-3:], oi3[-3:])
    for c, n in enumerate(sizes):
        for r in range(3):
            row = c * 3 + r
            ind = oi[row].numpy()
            if n == 0:
                assert (ind == -1).all() and float(x[row].abs().max()) == 0.0
                continue
            assert OC.resample_indices_ok(ind, n, N), (n, N)
            if n >= N:
                assert (np.diff(ind) > 0).all()                      # ascending = a subset
            src = pts[int(offsets[c]):int(offsets[c + 1])].numpy()
            assert np.array_equal(x[row].numpy(), src[ind].T)        # gathered, channel-major
    # different repeats draw different subsets
    assert not torch.equal(oi[-1], oi[-2])
    # without-replacement draws are uniform: every point of a 1500-set is picked with frequency ~ N/n
    if N == 500:
        freq = np.zeros(1500)
        xs, ois = prep.resample(pts, offsets, N, repeat=60, seed=7, return_index=True, lib=lib)
        for r in range(60):
            freq[ois[6 * 60 + r].numpy()] += 1
        assert abs(freq.mean() / 60 - N / 1500) < 1e-9 and freq.std() / 60 < 0.12
