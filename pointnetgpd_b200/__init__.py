"""pointnetgpd_b200 -- B200-native implementation of the PointNetGPD PointNet hot path.

    from pointnetgpd_b200.model.pointnet import PointNetCls     # reference call surface
    import pointnetgpd_b200; pointnetgpd_b200.install_as_model()  # make `import model.pointnet` resolve here

The arithmetic lives in libpgpd.so (include/pgpd.h), hand-written CUDA for sm_100a.
"""
import sys

__version__ = "0.1.0"


def install_as_model(force=False):
    """Register this package's `model` sub-package as the top-level `model` package, which is the name
    the reference scripts and pickled checkpoints use (`model.pointnet.PointNetCls`)."""
    from . import model as _model
    from .model import pointnet as _pointnet, gpd as _gpd
    if force or "model" not in sys.modules:
        sys.modules["model"] = _model
        sys.modules["model.pointnet"] = _pointnet
        sys.modules["model.gpd"] = _gpd
        try:
            from .model import dataset as _dataset
            sys.modules["model.dataset"] = _dataset
        except ImportError:
            pass
    return sys.modules["model"]
