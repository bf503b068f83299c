"""neural_compressor_b200 -- a B200-native (sm_100a) weight-only-quantisation engine behind
intel/neural-compressor's torch `prepare()/convert()/quantize()` API.

Host code is Python/PyTorch (plumbing: device memory, streams, torch.distributed); every hot-path
computation is a hand-written CUDA kernel in libb200woq.so reached through the C ABI declared in
include/b200woq.h.  There is no CPU fallback: ops raise if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
