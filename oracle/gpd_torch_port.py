"""CPU / eager-GPU port of the reference GPDClassifier (PointNetGPD/model/gpd.py:5-31), functional form.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker of pointnetgpd_b200.model.gpd.  Every arithmetic step of the
reference is a torch op; this restates the same sequence on a state dict (gpd.py line numbers in the comments) and is pinned
to the imported reference class in tests/test_oracle.py::test_gpd_port_matches_reference (build container only)."""
import torch
import torch.nn.functional as F


def gpd_forward(sd, x, dropout=False):
    h = F.max_pool2d(F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"]), 2, stride=2)      # gpd.py:22  pool1(conv1(x))
    h = F.max_pool2d(F.conv2d(h, sd["conv2.weight"], sd["conv2.bias"]), 2, stride=2)      # gpd.py:23  pool2(conv2(x))
    h = h.view(-1, 7200)                                                                  # gpd.py:24
    h = F.relu(F.linear(h, sd["fc1.weight"], sd["fc1.bias"]))                             # gpd.py:25
    assert not dropout                                                                    # gpd.py:26-27 (not used by any script)
    h = F.linear(h, sd["fc2.weight"], sd["fc2.bias"])                                     # gpd.py:28
    return F.log_softmax(h, dim=-1)                                                       # gpd.py:29


def make_gpd_state(seed, input_chann, dtype=torch.float32):
    """Deterministic torch-default-like initialisation from the package's splitmix64 generators."""
    from pointnetgpd_b200 import synth as W
    shapes = {"conv1.weight": (20, input_chann, 5, 5), "conv1.bias": (20,), "conv2.weight": (50, 20, 5, 5), "conv2.bias": (50,),
              "fc1.weight": (500, 7200), "fc1.bias": (500,), "fc2.weight": (2, 500), "fc2.bias": (2,)}
    fan = {"conv1": input_chann * 25, "conv2": 500, "fc1": 7200, "fc2": 500}
    sd = {}
    for i, (k, shp) in enumerate(shapes.items()):
        bound = 1.0 / (fan[k.split(".")[0]] ** 0.5)
        sd[k] = torch.tensor(W.uniform(seed * 131 + i, shp, -bound, bound), dtype=dtype)
    return sd
