"""Drop-in for the reference's models/correlation_native.py (:6-23), on our kernel."""
import torch.nn as nn

from .. import ops


class Correlation(nn.Module):
    def __init__(self, max_displacement=4, *args, **kwargs):
        super(Correlation, self).__init__()
        self.max_displacement = max_displacement
        self.output_dim = 2 * self.max_displacement + 1
        self.pad_size = self.max_displacement

    def forward(self, x1, x2):
        return ops.correlation(x1, x2, self.max_displacement)
