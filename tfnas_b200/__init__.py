"""tfnas_b200: B200-native TF-NAS supernet search path.

The north-star tolerance (1e-3 on logits and alpha-grads) needs true fp32 everywhere: SURVEY 8c measured a 2e-2 alpha-grad
error with single-pass TF32 operands.  The library's own GEMMs use the 3-term tf32 split; any torch op left around them
(a user's criterion, test references) must not silently drop to TF32 either -- in particular cuDNN/cuBLAS BACKWARD kernels,
which run inside loss.backward() long after a forward-time ``torch.backends.cudnn.flags`` scope has exited.
"""
import torch as _torch

_torch.backends.cudnn.allow_tf32 = False
_torch.backends.cuda.matmul.allow_tf32 = False
