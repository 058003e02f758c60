"""Synthetic `$PointNetGPD_FOLDER` tree in the reference's on-disk formats:

  PointNetGPD/data/google2cloud.pkl                      dict: object -> (cloud object name, float64 [4,4])   dataset.py:13
  PointNetGPD/data/ycb_grasp/{train,test}/<obj>.npy      float [rows,12]: center3, axis3, width, angle, jaw_width,
                                                         min_width, friction score, canny score          (grasp.py:235-246 +
                                                         generate-dataset-canny.py:47-53)
  data/ycb-tools/models/ycb/<obj>/rgbd/clouds/pc_NP3_NP5_<i>.npy   float32 [P,3] view clouds             dataset.py:400

Grasps are placed so that their gripper box contains enough cloud points (>= 50, dataset.py:71-72) and their scores are
on either side of the 0.6 thresholds (main_1v.py:54-55).

The module also holds the deterministic, library-independent generators of model state, grasp clouds and labels
(`make_state`, `make_clouds`, `make_labels`, built on a splitmix64 counter generator instead of torch's RNG so that
the same seeds give the same arrays in the build container and on the GPU box).  bench.py, smoke() and the tests use
them for synthetic inputs; oracle/weights.py re-exports them for the golden-vector scripts.

State-dict key names / shapes follow the reference classes (/root/reference/PointNetGPD/model/pointnet.py:8-25 STN3d,
:123-135 PointNetfeat, :177-187 PointNetCls): 74 entries for PointNetCls."""
import os
import pickle

import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(idx):
    """Vectorised splitmix64 of uint64 counters -> uint64."""
    with np.errstate(over="ignore"):
        z = (idx + np.uint64(0x9E3779B97F4A7C15)) & _MASK
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
        return z ^ (z >> np.uint64(31))


def uniform01(seed, n):
    """n doubles in [0,1), a pure function of (seed, position)."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([seed], dtype=np.uint64))[0]
        idx = np.arange(n, dtype=np.uint64) + base
    bits = _splitmix64(idx) >> np.uint64(11)
    return bits.astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform(seed, shape, lo, hi):
    n = int(np.prod(shape)) if len(shape) else 1
    return (lo + (hi - lo) * uniform01(seed, n)).reshape(shape)


def normal(seed, shape):
    n = int(np.prod(shape)) if len(shape) else 1
    u1 = uniform01(seed, n)
    u2 = uniform01(seed ^ 0x5DEECE66D, n)
    return (np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)).reshape(shape)


def _stable_hash(name):
    h = 1469598103934665603
    for ch in name.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


# (prefix, [(conv/fc name, out, in, is_conv)], [(bn name, channels)])
def _layout(k):
    return [
        ("feat.stn.", [("conv1", 64, 3, True), ("conv2", 128, 64, True), ("conv3", 1024, 128, True),
                       ("fc1", 512, 1024, False), ("fc2", 256, 512, False), ("fc3", 9, 256, False)],
         [("bn1", 64), ("bn2", 128), ("bn3", 1024), ("bn4", 512), ("bn5", 256)]),
        ("feat.", [("conv1", 64, 3, True), ("conv2", 128, 64, True), ("conv3", 1024, 128, True)],
         [("bn1", 64), ("bn2", 128), ("bn3", 1024)]),
        ("", [("fc1", 512, 1024, False), ("fc2", 256, 512, False), ("fc3", k, 256, False)],
         [("bn1", 512), ("bn2", 256)]),
    ]


def _dual_layout(k):
    """DualPointNetCls(input_chann=6) (pointnet.py:157-174): feat.stn1 / feat.stn2 (SimpleSTN3d, :48-85), the 6-channel trunk
    (DualPointNetfeat, :88-120) and the classifier head."""
    def stn(p):
        return (p, [("conv1", 64, 3, True), ("conv2", 128, 64, True), ("conv3", 256, 128, True),
                    ("fc1", 128, 256, False), ("fc2", 64, 128, False), ("fc3", 9, 64, False)],
                [("bn1", 64), ("bn2", 128), ("bn3", 256), ("bn4", 128), ("bn5", 64)])
    return [stn("feat.stn1."), stn("feat.stn2."),
            ("feat.", [("conv1", 64, 6, True), ("conv2", 128, 64, True), ("conv3", 1024, 128, True)],
             [("bn1", 64), ("bn2", 128), ("bn3", 1024)]),
            ("", [("fc1", 512, 1024, False), ("fc2", 256, 512, False), ("fc3", k, 256, False)],
             [("bn1", 512), ("bn2", 256)])]


def state_keys(k=2, dual=False):
    """The 74 state_dict keys of the reference PointNetCls (dual: the 111 of DualPointNetCls), in registration order
    (conv/fc first, then bn, per module -- pointnet.py:11-25,126-132,181-186)."""
    keys = []
    # registration order inside each reference __init__: convs, (mp1), fcs, relu, bns
    # but nested: PointNetCls registers feat (PointNetfeat: stn (STN3d), conv*, bn*), fc*, bn*
    for prefix, lins, bns in (_dual_layout(k) if dual else _layout(k)):
        # feat.stn.* comes before feat.conv* because stn is the first attribute set
        for name, _o, _i, _c in lins:
            keys += [prefix + name + ".weight", prefix + name + ".bias"]
        for name, _c in bns:
            keys += [prefix + name + s for s in
                     (".weight", ".bias", ".running_mean", ".running_var", ".num_batches_tracked")]
    return keys


def make_state(seed, k=2, style="default", dtype=np.float32, dual=False):
    """Build a full PointNetCls (dual: DualPointNetCls(input_chann=6)) state dict as numpy arrays.

    style="default": torch default-init distributions (Conv1d/Linear U(+-1/sqrt(fan_in)),
        BN gamma=1 beta=0 rm=0 rv=1)  -- pointnet.py:178-187 default constructors.
    style="wild": random BN affine incl. NEGATIVE gammas, non-trivial running stats
        (the shipped checkpoint has 143/1024 negative gammas in feat.bn3 -- SURVEY App. B).
    """
    sd = {}
    for prefix, lins, bns in (_dual_layout(k) if dual else _layout(k)):
        for name, o, i, is_conv in lins:
            bound = 1.0 / np.sqrt(i)
            w = uniform(seed ^ _stable_hash(prefix + name + ".weight"), (o, i), -bound, bound)
            b = uniform(seed ^ _stable_hash(prefix + name + ".bias"), (o,), -bound, bound)
            if style == "wild":
                w = w * 1.7
            sd[prefix + name + ".weight"] = (w.reshape(o, i, 1) if is_conv else w).astype(dtype)
            sd[prefix + name + ".bias"] = b.astype(dtype)
        for name, c in bns:
            if style == "wild":
                g = uniform(seed ^ _stable_hash(prefix + name + ".g"), (c,), 0.5, 1.5)
                sgn = np.where(uniform01(seed ^ _stable_hash(prefix + name + ".s"), c) < 0.2, -1.0, 1.0)
                g = g * sgn
                be = uniform(seed ^ _stable_hash(prefix + name + ".b"), (c,), -0.5, 0.5)
                rm = 0.5 * normal(seed ^ _stable_hash(prefix + name + ".rm"), (c,))
                rv = uniform(seed ^ _stable_hash(prefix + name + ".rv"), (c,), 0.5, 2.0)
            else:
                g, be, rm, rv = np.ones(c), np.zeros(c), np.zeros(c), np.ones(c)
            sd[prefix + name + ".weight"] = g.astype(dtype)
            sd[prefix + name + ".bias"] = be.astype(dtype)
            sd[prefix + name + ".running_mean"] = rm.astype(dtype)
            sd[prefix + name + ".running_var"] = rv.astype(dtype)
            sd[prefix + name + ".num_batches_tracked"] = np.array(0, dtype=np.int64)
    return sd


GRIPPER_W = 0.085  # robotiq_85 max_width; dataset.py:57-59 crop box is (w/4, w/2, w/4)


def make_clouds(seed, B, N, kind="box", dtype=np.float32):
    """Synthetic grasp clouds [B,3,N] (channel-major, the model's input layout,
    main_1v.py:69; values per SURVEY section 8d).

    kind="box":   uniform in the gripper crop box (|x|<w/4, |y|<w/2, |z|<w/4).
    kind="dup":   first N//3 points unique, the rest resampled WITH replacement
                  from them (mimics dataset.py:439-444 when a crop has < N points).
    kind="randn": standard normal (conditioning check, not physical).
    """
    if kind == "randn":
        return normal(seed, (B, 3, N)).astype(dtype)
    half = np.array([GRIPPER_W / 4, GRIPPER_W / 2, GRIPPER_W / 4]).reshape(1, 3, 1)
    x = uniform(seed, (B, 3, N), -1.0, 1.0) * half
    if kind == "dup":
        nu = max(1, N // 3)
        pick = (uniform01(seed ^ 0xABCDEF, B * (N - nu)).reshape(B, N - nu) * nu).astype(np.int64)
        for b in range(B):
            x[b, :, nu:] = x[b][:, pick[b]]
    elif kind != "box":
        raise ValueError(kind)
    return x.astype(dtype)


def make_labels(seed, B, k):
    return (uniform01(seed ^ 0x1234567, B) * k).astype(np.int64)


# ---- synthetic $PointNetGPD_FOLDER tree --------------------------------------------------------------------------------
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
