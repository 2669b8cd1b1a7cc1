"""starr_b200 -- the SumTreeArray hot path of justinmaojones/starr on NVIDIA B200 (sm_100a).

Exports the same six names as the reference package (``starr/__init__.py:1-6``) plus the
batched device entry point ``DeviceSumTree`` (torch tensors in, torch tensors out, no host
sync) and ``ShardedSumTree`` (leaf range sharded by subtree over 2/4/8 GPUs).

There is no CPU implementation in this package: every call runs CUDA kernels from
``libstarr_b200.so`` and fails loudly if the library or a GPU is missing.
"""
from .kernels import (build_sumtree_from_array, get_prefix_sum_idx, strided_sum, sum_over,
                      update_sumtree)
from .sumtree_array import SumTreeArray

__all__ = ["get_prefix_sum_idx", "strided_sum", "sum_over", "build_sumtree_from_array",
           "update_sumtree", "SumTreeArray", "DeviceSumTree", "DeviceMinTree", "ShardedSumTree"]


def __getattr__(name):
    # torch is imported only when the device entry points are used
    if name == "DeviceSumTree":
        from .device_tree import DeviceSumTree
        return DeviceSumTree
    if name == "DeviceMinTree":
        from .device_tree import DeviceMinTree
        return DeviceMinTree
    if name == "ShardedSumTree":
        from .sharded import ShardedSumTree
        return ShardedSumTree
    raise AttributeError(name)
