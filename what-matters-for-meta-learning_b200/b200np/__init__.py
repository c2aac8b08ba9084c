"""b200np -- host side of the B200-native neural-process hot path.

``lib``     ctypes binding of csrc/libb200np.so (the C ABI in include/b200np.h)
``ops``     tensor-level wrappers (allocation + launch on torch's current stream)
``engine``  autograd segments and the model-family forward passes
``optim``   FusedAdam over flat parameter / gradient buffers
``dist``    task sharding + NCCL gradient all-reduce (one process per GPU)
"""
