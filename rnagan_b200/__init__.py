"""rnagan_b200 -- B200-native (sm_100a) implementation of the RNA-GAN training + synthesis hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every heavy operation is a hand-written
CUDA kernel in ``csrc/`` reached through the C ABI of ``include/rnagan_b200.h``.  There is no CPU fallback.
"""
__version__ = "0.1.0"
