"""Drop-in replacement of the reference's `model` package (PointNetGPD/model/): same module
names, class names, constructor signatures, sub-module / parameter names and return values; the
PointNet forward/backward runs in libpgpd (hand-written sm_100a CUDA) instead of torch.nn ops."""
