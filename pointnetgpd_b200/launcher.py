"""Run the reference's experiment scripts UNCHANGED against this package.

    python -m pointnetgpd_b200.launcher /path/to/PointNetGPD/main_1v.py --mode train --batch-size 64 --cuda --gpu 0
    python -m pointnetgpd_b200.launcher --synthetic-data /tmp/pgpd_data /path/to/PointNetGPD/main_1v.py ...

What it fabricates around the byte-identical script:
  * `model`, `model.pointnet`, `model.dataset`, `model.gpd` resolve to this package (`install_as_model`), so the
    script's `from model.pointnet import PointNetCls, DualPointNetCls` builds the libpgpd-backed classes;
  * `tensorboardX.SummaryWriter` (main_1v.py:12,40) -> torch.utils.tensorboard if importable, else a no-op logger;
  * `torch.load` of whole-module pickles: `weights_only=False` default (main_1v.py:153) and the
    `torch.nn.backends.thnn` stub the 2018 checkpoint needs;
  * `scipy.stats.mode` accepts the CUDA tensors main_test.py:92 hands it under --cuda (host copies);
  * `./assets/learned_models` exists (the *_mc / fullv scripts never create it);
  * optionally a synthetic `$PointNetGPD_FOLDER` tree (`--synthetic-data DIR`, see `synth.py`).
The script itself is executed with `runpy.run_path(..., run_name="__main__")`, which does not put the script's own
directory on `sys.path`, so the reference's `model/` package next to it is never imported.
"""
import os
import runpy
import sys
import types


class _NullWriter:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def close(self):
        pass


def install_shims():
    import pointnetgpd_b200
    pointnetgpd_b200.install_as_model(force=True)
    if "tensorboardX" not in sys.modules:
        tbx = types.ModuleType("tensorboardX")
        try:
            from torch.utils.tensorboard import SummaryWriter     # needs the tensorboard package
            tbx.SummaryWriter = SummaryWriter
        except Exception:
            tbx.SummaryWriter = _NullWriter
        sys.modules["tensorboardX"] = tbx
    if "torch.nn.backends.thnn" not in sys.modules:
        thnn = types.ModuleType("torch.nn.backends.thnn")
        thnn._get_thnn_function_backend = lambda: None
        sys.modules["torch.nn.backends.thnn"] = thnn
    try:
        # main_test.py:92 calls scipy.stats.mode on a list of tensors; with --cuda they are CUDA tensors, which numpy
        # cannot read (TypeError in the unmodified reference as well).  Hand scipy host copies.
        import scipy.stats as _st
        import torch as _torch
        if not getattr(_st.mode, "_pgpd_shim", False):
            _orig_mode = _st.mode

            def _mode(a, *args, **kwargs):
                if isinstance(a, (list, tuple)):
                    a = [t.detach().cpu().numpy() if isinstance(t, _torch.Tensor) else t for t in a]
                elif isinstance(a, _torch.Tensor):
                    a = a.detach().cpu().numpy()
                return _orig_mode(a, *args, **kwargs)
            _mode._pgpd_shim = True
            _st.mode = _mode
    except Exception:
        pass
    os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")
    os.makedirs("./assets/learned_models", exist_ok=True)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if argv and argv[0] == "--synthetic-data":
        from .synth import make_tree
        root = argv[1]
        make_tree(root)
        os.environ["PointNetGPD_FOLDER"] = root
        argv = argv[2:]
    if not argv:
        raise SystemExit(__doc__)
    script, script_args = argv[0], argv[1:]
    install_shims()
    sys.argv = [script] + script_args
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
