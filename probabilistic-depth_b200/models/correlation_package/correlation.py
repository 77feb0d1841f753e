"""Drop-in for the reference's models/correlation_package/correlation.py (:46-61).

The reference module wraps a CUDA extension (correlation_cuda.forward,
correlation_cuda.cc:10-87) built for sm_50..61.  Same constructor; forward only
(the north-star path is eval).  Supported configuration = the one the reference
instantiates (models/pwclite.py:123-125): kernel_size 1, both strides 1,
pad_size == max_displacement, corr_multiply 1.
"""
from torch.nn.modules.module import Module

from ... import ops


class Correlation(Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2,
                 corr_multiply=1):
        super(Correlation, self).__init__()
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply

    def forward(self, input1, input2):
        if not (self.kernel_size == 1 and self.stride1 == 1 and self.stride2 == 1 and
                self.pad_size == self.max_displacement and self.corr_multiply == 1):
            raise NotImplementedError(
                "Correlation: only kernel_size=1, stride1=stride2=1, pad_size=max_displacement, "
                "corr_multiply=1 has an sm_100a kernel")
        return ops.correlation(input1, input2, self.max_displacement)
