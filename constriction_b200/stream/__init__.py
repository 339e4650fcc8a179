"""Mirror of `constriction.stream` (reference: src/pybindings/stream/mod.rs:50-58): one coder per
object, same class names, method names, argument meaning and exceptions.  Every symbol is coded by
the batched CUDA kernels (a batch of one stream, coder state carried between calls through the C
ABI's raw-state interface), so these classes are bit-identical to the batch API by construction and
are what the reference's golden vectors are replayed against on the GPU.  They are the
compatibility surface, not the fast path: use `constriction_b200.batch` for throughput."""
from . import model, queue, stack  # noqa: F401
from .model import (Bernoulli, Binomial, Categorical, CustomModel, QuantizedCauchy, QuantizedGaussian,  # noqa: F401
                    QuantizedLaplace, ScipyModel, Uniform)
from .queue import RangeDecoder, RangeEncoder  # noqa: F401
from .stack import AnsCoder  # noqa: F401
