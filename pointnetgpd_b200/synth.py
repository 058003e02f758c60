"""Synthetic `$PointNetGPD_FOLDER` tree in the reference's on-disk formats:

  PointNetGPD/data/google2cloud.pkl                      dict: object -> (cloud object name, float64 [4,4])   dataset.py:13
  PointNetGPD/data/ycb_grasp/{train,test}/<obj>.npy      float [rows,12]: center3, axis3, width, angle, jaw_width,
                                                         min_width, friction score, canny score          (grasp.py:235-246 +
                                                         generate-dataset-canny.py:47-53)
  data/ycb-tools/models/ycb/<obj>/rgbd/clouds/pc_NP3_NP5_<i>.npy   float32 [P,3] view clouds             dataset.py:400

Grasps are placed so that their gripper box contains enough cloud points (>= 50, dataset.py:71-72) and their scores are
on either side of the 0.6 thresholds (main_1v.py:54-55)."""
import os
import pickle

import numpy as np


def make_tree(root, objects=("003_cracker_box",), train_rows=6500, test_rows=500, views=3, points=8000, seed=0):
    rng = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "PointNetGPD", "data"), exist_ok=True)
    transforms = {}
    for obj in objects:
        transforms[obj] = (obj, np.eye(4))
        cdir = os.path.join(root, "data", "ycb-tools", "models", "ycb", obj, "rgbd", "clouds")
        os.makedirs(cdir, exist_ok=True)
        for v in range(views):
            np.save(os.path.join(cdir, "pc_NP3_NP5_%d.npy" % v), rng.uniform(-0.05, 0.05, size=(points, 3)).astype(np.float32))
        for tag, rows in (("train", train_rows), ("test", test_rows)):
            gdir = os.path.join(root, "PointNetGPD", "data", "ycb_grasp", tag)
            os.makedirs(gdir, exist_ok=True)
            g = np.zeros((rows, 12))
            g[:, 0:3] = rng.uniform(-0.01, 0.01, size=(rows, 3))
            ax = rng.normal(size=(rows, 3))
            g[:, 3:6] = ax / np.linalg.norm(ax, axis=1, keepdims=True)
            g[:, 6] = 0.085
            g[:, 7] = rng.uniform(-np.pi, np.pi, size=rows)
            g[:, 8] = 0.085
            g[:, 9] = 0.0
            g[:, 10] = np.where(rng.rand(rows) < 0.5, 0.4, 1.2)        # friction score: <= 0.6 good, >= 0.6 bad
            g[:, 11] = rng.uniform(0, 1, size=rows) * 0.001
            np.save(os.path.join(gdir, obj + ".npy"), g)
    with open(os.path.join(root, "PointNetGPD", "data", "google2cloud.pkl"), "wb") as f:
        pickle.dump(transforms, f)
    return root
