"""Batched candidate scoring: what the reference's deployment loop does one grasp at a time
(dex-net/apps/kinect2grasp.py:454-491 calling main_test.test_network, main_test.py:59-69), as ONE batched forward.

    pred, score = score_candidates(model, clouds, input_points_num=500)

For every candidate (its points already in the hand frame, `collect_pc` of kinect2grasp.py:238-258):
  * fewer than `min_points` points -> prediction 0, score 0.0                     (kinect2grasp.py:462-466)
  * `repeat` times: resample to `input_points_num` points -- without replacement if it has enough points, with
    replacement otherwise (:473-478) -- on the GPU (libpgpd `pgpd_resample`);
  * one eval-mode, no-grad forward over all (candidate, repeat) rows; softmax; arg-max   (main_test.py:65-69)
  * majority vote over the repeats; score = mean probability of the best class (last column) over the repeats that
    agree with the vote (:483-488).
"""
import numpy as np
import torch

from . import prep


def score_candidates(model, clouds, input_points_num=None, min_points=20, repeat=1, seed=0, max_batch=8192):
    """clouds: list of [n_i,3] arrays/tensors.  Returns (pred: int64 [C], score: float64 [C]) as numpy arrays."""
    dev = next(model.parameters()).device
    N = input_points_num or model.num_points
    C_ = len(clouds)
    pred = np.zeros(C_, dtype=np.int64)
    score = np.zeros(C_, dtype=np.float64)
    keep = [i for i, c in enumerate(clouds) if len(c) >= min_points]
    if not keep:
        return pred, score
    lens = [len(clouds[i]) for i in keep]
    offsets = torch.zeros(len(keep) + 1, dtype=torch.int32)
    offsets[1:] = torch.cumsum(torch.tensor(lens, dtype=torch.int64), 0).to(torch.int32)
    cat = torch.cat([torch.as_tensor(np.asarray(clouds[i]), dtype=torch.float32) for i in keep], 0).to(dev)
    x = prep.resample(cat, offsets, N, repeat=repeat, seed=seed)            # [K*repeat, 3, N]
    was_training = model.training
    model.eval()
    probs = []
    with torch.no_grad():
        for s in range(0, x.shape[0], max_batch):
            logp, _ = model(x[s:s + max_batch])
            probs.append(logp.exp())                                         # softmax = exp(log_softmax)
    if was_training:
        model.train()
    probs = torch.cat(probs, 0).double().cpu().numpy().reshape(len(keep), repeat, -1)
    best = probs.shape[2] - 1                                                # 3-class: column 2, 2-class: column 1
    p = probs.argmax(2)
    for j, i in enumerate(keep):
        vals, counts = np.unique(p[j], return_counts=True)
        v = vals[np.argmax(counts)]                                          # scipy.stats.mode: smallest most common
        pred[i] = v
        score[i] = probs[j][p[j] == v][:, best].mean()
    return pred, score
