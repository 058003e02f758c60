"""`model.gpd` -- the GPD baseline CNN (PointNetGPD/model/gpd.py:5-31) is a different network
(LeNet on 60x60 projections), not on the PointNet hot path and not named by the metric; it is
outside this package's scope.  The name resolves so `from model.gpd import *` works."""
import torch.nn as nn

__all__ = ["GPDClassifier"]


class GPDClassifier(nn.Module):
    def __init__(self, input_chann, dropout=False):
        raise NotImplementedError("GPDClassifier (model/gpd.py:5-31) is the paper's baseline CNN and is outside "
                                  "the scope of pointnetgpd_b200 (SURVEY.md section 8f, rank 4)")
