"""numpy restatement of the steps immediately BEFORE the model.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  * gripper-box crop      PointNetGPD/model/dataset.py:15-76   (BaseGraspDataset.collect_pc, projection=False)
  * resample to N points  PointNetGPD/model/dataset.py:438-444 / dex-net/apps/kinect2grasp.py:473-478
  * candidate scoring     dex-net/apps/kinect2grasp.py:454-491 (+ main_test.py:59-69 test_network)

Written independently from the product's host code (pointnetgpd_b200/prep.py) so the two can be compared.
"""
import numpy as np


def grasp_frame(grasp, transform):
    """center (3,), rotation matrix (3,3) with rows (approach, binormal, minor_normal), width.
    dataset.py:16-51: grasp = [center3, axis3, width, angle, ...], transform = 4x4 mesh->cloud."""
    c = np.asarray(grasp[0:3], dtype=np.float64)
    ax = np.asarray(grasp[3:6], dtype=np.float64)
    width, angle = float(grasp[6]), float(grasp[7])
    ax = ax / np.linalg.norm(ax)                                    # :21
    ct, st = np.cos(angle), np.sin(angle)                           # :24-25
    R1 = np.array([[ct, 0, -st], [0, 1, 0], [st, 0, ct]])           # np.c_ of the three columns, :26
    ay = ax
    axx = np.array([ay[1], -ay[0], 0.0])                            # :28
    if np.linalg.norm(axx) == 0:                                    # :29-30
        axx = np.array([1.0, 0.0, 0.0])
    axx = axx / np.linalg.norm(axx)
    ay = ay / np.linalg.norm(ay)
    az = np.cross(axx, ay)                                          # :33
    R2 = np.stack([axx, ay, az], axis=1)                            # columns, :34
    approach = R2.dot(R1)[:, 0]                                     # :35
    approach = approach / np.linalg.norm(approach)
    minor = np.cross(ax, approach)                                  # :37
    T = np.asarray(transform, dtype=np.float64)
    center = T.dot(np.append(c, 1.0))[:3]                           # :45
    binormal = T.dot(np.append(ax, 0.0))[:3]                        # :46
    approach = T.dot(np.append(approach, 0.0))[:3]                  # :47
    minor = T.dot(np.append(minor, 0.0))[:3]                        # :48-49
    return center, np.stack([approach, binormal, minor], axis=0), width   # :50


def crop(pc, grasp, transform):
    """indices (ascending) of the cloud points inside the gripper box, and their local coordinates (float64).
    dataset.py:51-76."""
    center, M, width = grasp_frame(grasp, transform)
    pc_t = (M.dot((np.asarray(pc) - center).T)).T                   # :53
    xl, yl, zl = width / 4, width / 2, width / 4                    # :57-59
    inside = ((pc_t[:, 0] > -xl) & (pc_t[:, 0] < xl) & (pc_t[:, 1] > -yl) & (pc_t[:, 1] < yl)
              & (pc_t[:, 2] > -zl) & (pc_t[:, 2] < zl))             # :61-69
    idx = np.where(inside)[0]
    return idx, pc_t[idx]


def resample_indices_ok(idx, n, N):
    """Property of the reference's resampling (dataset.py:439-444, kinect2grasp.py:473-478): N indices into a set of
    n points, all distinct when n >= N (np.random.choice(..., replace=False)), any valid index otherwise."""
    idx = np.asarray(idx)
    if idx.shape != (N,) or idx.min() < 0 or idx.max() >= n:
        return False
    return len(np.unique(idx)) == N if n >= N else True


def vote(pred_rows, prob_rows, best_col):
    """kinect2grasp.py:483-491: majority vote over the `repeat` predictions (scipy.stats.mode: smallest of the most
    common values) and mean probability of the best class over the repeats that agree with the vote."""
    pred_rows = np.asarray(pred_rows)
    vals, counts = np.unique(pred_rows, return_counts=True)
    v = vals[np.argmax(counts)]
    sel = pred_rows == v
    return int(v), float(np.mean(np.asarray(prob_rows)[sel][:, best_col]))
