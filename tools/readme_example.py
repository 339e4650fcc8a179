"""Runs the README's usage examples (GPU)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from constriction_b200 import batch as B
from constriction_b200 import stream as constriction_stream

message = np.array([6, 10, -4, 2, 5, 2, 1, 0, 2], dtype=np.int32)
means = np.array([2.5, 13.1, -1.1, -3.0, 6.1, 2.4, 0.3, 0.1, 1.9])
stds = np.array([4.1, 8.7, 6.2, 5.4, 24.1, 12.8, 4.9, 28.9, 4.2])
family = constriction_stream.model.QuantizedGaussian(-100, 100)
coder = constriction_stream.stack.AnsCoder()
coder.encode_reverse(message, family, means, stds)
words = coder.get_compressed()
assert np.all(constriction_stream.stack.AnsCoder(words).decode(family, means, stds) == message)
print("mirror:", words)

bc = B.BatchCoder()
model = B.ModelTable.quantized_gaussian(-50, 50, [3.2], [9.6])
symbols = torch.randint(-50, 51, (100_000_000,), dtype=torch.int32, device="cuda")
comp = bc.ans_encode(symbols, model, n_streams=148 * 1024)
decoded = bc.ans_decode(comp, model)
bc.check()
assert torch.equal(decoded, symbols)

offsets = torch.arange(0, 801, device="cuda") * 122_070
latents = symbols[: 800 * 122_070]
comp = bc.range_encode(latents, model, sym_offsets=offsets, checkpoint_every=1024)
decoded = bc.range_decode(comp, model)
bc.check()
assert torch.equal(decoded, latents)

n = latents.numel()
means_dev = torch.randn(n, device="cuda", dtype=torch.float64) * 3
stds_dev = torch.rand(n, device="cuda", dtype=torch.float64) * 5 + 0.5
params = B.GaussianParams(-64, 64, means_dev, stds_dev)
comp = bc.ans_encode(latents, params, sym_offsets=offsets)
bc.check()
assert torch.equal(bc.ans_decode(comp, params), latents)
print("README examples OK")
