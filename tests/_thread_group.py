"""TEST HELPER ONLY: a stand-in for ``torch.distributed`` that runs the ranks of a ShardedSumTree
as threads of ONE process sharing ONE GPU, so the complete CUDA path of the sharded tree (plan,
slab exchange, top fold, routing) is exercised on a single-GPU box.  Collectives are implemented
with a barrier and device copies on the shared default stream."""
import threading

import torch


class ThreadGroup:
    def __init__(self, world: int):
        self.world = world
        self._barrier = threading.Barrier(world)
        self._slots = [None] * world
        self._local = threading.local()

    # -- the subset of torch.distributed that starr_b200/sharded.py uses ------------------------
    def bind(self, rank: int) -> None:
        self._local.rank = rank

    def is_initialized(self) -> bool:
        return True

    def get_world_size(self, group=None) -> int:
        return self.world

    def get_rank(self, group=None) -> int:
        return self._local.rank

    def barrier(self, group=None) -> None:
        self._barrier.wait()

    def all_gather_into_tensor(self, output: torch.Tensor, input: torch.Tensor, group=None) -> None:
        self._slots[self._local.rank] = input
        torch.cuda.current_stream().synchronize()  # ranks may run on streams of their own: the input is complete
        self._barrier.wait()
        k = input.numel()
        flat = output.view(-1)
        for r in range(self.world):
            flat[r * k:(r + 1) * k].copy_(self._slots[r].reshape(-1))
        torch.cuda.current_stream().synchronize()
        self._barrier.wait()                      # nobody overwrites its slot before all have copied

    def all_gather_object(self, object_list: list, obj, group=None) -> None:
        self._slots[self._local.rank] = obj
        self._barrier.wait()
        object_list[:] = list(self._slots)
        self._barrier.wait()

    def run(self, fn):
        """fn(rank) on every rank; returns the list of results, re-raises the first exception."""
        results, errors = [None] * self.world, []

        def body(r):
            self.bind(r)
            try:
                results[r] = fn(r)
            except BaseException as e:            # noqa: BLE001 - reported to the main thread
                errors.append(e)
                self._barrier.abort()

        threads = [threading.Thread(target=body, args=(r,)) for r in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return results
