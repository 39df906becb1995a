"""On-GPU replay store behind the reference's `next(replay_iter)` contract (reference tools/replay.py).

The reference keeps episodes as host numpy arrays and builds every batch with B Python-level slices per key, `np.stack`, then a
pageable host->device copy of ~30 MB (tools/replay.py:223-236).  Once the update itself takes ~20 ms that loop is the bottleneck.
Here every key of every stored episode lives in ONE flat device buffer [capacity_steps, row_bytes] (HBM: 12.3 KB per 64x64x3 frame,
a 5 M-step buffer is 62 GB of the B200's 180 GB); `next(iter(buffer))` samples the (episode, offset) pairs on the host with the SAME
numpy calls as the reference — so a given `np.random.seed` yields the same windows — uploads B int64 start rows and launches one
gather kernel per key (csrc/replay.cu).  The yielded dict has the reference's keys, dtypes and [B, L, ...] shapes on `device`.

Mirrored surface: ReplayBuffer(data_specs, meta_specs, directory, length, capacity, minlen, min_t_sampling, device, ...),
.add_episode / .store_episode / .stats / __iter__ / __len__, make_replay_loader(buffer, batch_size).  Saving/loading .npz episode
files, the per-step `add()` of the env loop and the DataLoader wrapper are host-side storage code and stay with the reference
(out of scope); episodes arrive here as dicts of numpy arrays exactly as `store_episode(episode=...)` receives them.
"""
import collections

import numpy as np
import torch

from ._lib import lib, GenrlError


def eplen(episode):
    return len(episode["action"])


def convert(value):
    """reference tools/replay.py:297-305: floats -> float32, signed ints -> int32, uint8 stays"""
    value = np.asarray(value)
    if np.issubdtype(value.dtype, np.floating):
        return value.astype(np.float32)
    if np.issubdtype(value.dtype, np.signedinteger):
        return value.astype(np.int32)
    return value


_TORCH_DT = {np.dtype("float32"): torch.float32, np.dtype("uint8"): torch.uint8, np.dtype("bool"): torch.bool, np.dtype("int32"): torch.int32,
             np.dtype("int64"): torch.int64, np.dtype("float64"): torch.float64}


class ReplayBuffer:
    def __init__(self, data_specs=None, meta_specs=None, directory=None, length=20, capacity=0, ongoing=False, minlen=1, maxlen=0,
                 prioritize_ends=False, device="cuda", load_first=False, save_episodes=False, ignore_extra_keys=False,
                 load_recursive=False, min_t_sampling=0, keys=None, max_episode_steps=1024, **kwargs):
        if capacity <= 0:
            raise GenrlError("the device replay store needs an explicit capacity in steps")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise GenrlError("genrl_b200.replay.ReplayBuffer keeps episodes in device memory: device must be a CUDA device (no CPU fallback)")
        self._length = length
        self._capacity = int(capacity)
        # Episodes are stored contiguously (a window gather is then B plain spans), so the ring carries `max_episode_steps` rows of slack:
        # with FIFO placement the only hole is the skipped tail at the physical end (< one episode), and capacity + one episode of rows is
        # enough for the reference's eviction rule (drop the oldest while loaded_steps + length > capacity) to be the ONLY eviction.
        self._rows = self._capacity + int(max(max_episode_steps, maxlen or 0))
        self._minlen = minlen
        self._min_t_sampling = min_t_sampling
        self._ignore_extra_keys = ignore_extra_keys
        self._keys = list(keys) if keys is not None else None        # fixed at the first episode otherwise
        self._store = {}                                             # key -> device uint8 [capacity, row_bytes]
        self._meta = {}                                              # key -> (numpy dtype, row shape)
        self._eps = collections.deque()                              # (first_row, length) of the stored episodes, oldest first
        self._episode_lens = np.array([])
        self._head = 0                                               # next free row
        self._total_steps = self._total_episodes = self._loaded_steps = self._loaded_episodes = 0
        self.batch_size = None
        self._starts_host = None
        self._starts_dev = None

    # ---- bookkeeping (reference :112-119, :280-282) ----
    @property
    def stats(self):
        return {"total_steps": self._total_steps, "total_episodes": self._total_episodes, "loaded_steps": self._loaded_steps,
                "loaded_episodes": self._loaded_episodes}

    def __len__(self):
        return self._total_steps

    # ---- storing ----
    def add_episode(self, episode):
        """reference :163-173 (without the .npz save: file storage stays with the reference)"""
        length = eplen(episode)
        if length < self._minlen:
            return False
        self._total_steps += length
        self._total_episodes += 1
        return self.store_episode(episode={k: convert(v) for k, v in episode.items()})

    def _alloc(self, episode):
        for k, v in episode.items():
            if self._keys is not None and k not in self._keys:
                continue
            v = np.asarray(v)
            row = int(np.prod(v.shape[1:], dtype=np.int64)) * v.dtype.itemsize
            self._meta[k] = (v.dtype, tuple(v.shape[1:]))
            self._store[k] = torch.empty(self._rows, max(row, 1), device=self.device, dtype=torch.uint8)
        self._keys = list(self._store)

    def _evict_oldest(self):
        first, n = self._eps.popleft()
        self._episode_lens = self._episode_lens[1:]
        self._loaded_steps -= n
        self._loaded_episodes -= 1

    def store_episode(self, filename=None, episode=None, run_checks=True):
        """reference :175-221: append an episode, dropping the oldest ones while the step capacity would be exceeded"""
        if filename is not None:
            raise GenrlError("loading episode files is the reference's storage code: pass episode=dict of arrays")
        if not episode:
            return False
        length = eplen(episode)
        if length > self._capacity:
            raise GenrlError(f"episode of {length} steps exceeds the store's capacity of {self._capacity}")
        if length < self._length:
            pass        # stored like the reference does; sampling it raises in np.random.randint (low >= high), as in the reference
        if not self._store:
            self._alloc(episode)
        for k in episode:
            if k not in self._store:
                if self._ignore_extra_keys:
                    continue
                raise KeyError("Extra key ", k)
            if run_checks:
                v = np.asarray(episode[k])
                assert len(v) == length, f"Found {v.shape} VS eplen: {length}"
                assert (v.dtype, tuple(v.shape[1:])) == self._meta[k], f"for ({k}) expected {self._meta[k]}, received {(v.dtype, v.shape[1:])}"
        while self._loaded_steps + length > self._capacity:                  # enforce limit (reference :201-207)
            self._evict_oldest()
        # contiguous placement in the ring: an episode never wraps
        if self._head + length > self._rows:
            self._head = 0
        lo, hi = self._head, self._head + length
        while self._eps and any(f < hi and lo < f + n for f, n in self._eps):
            self._evict_oldest()            # only reachable when an episode is longer than the slack (max_episode_steps)
        for k in self._store:
            v = np.ascontiguousarray(np.asarray(episode[k]))
            src = torch.from_numpy(v.reshape(length, -1).view(np.uint8) if v.dtype != np.bool_ else v.reshape(length, -1).astype(np.uint8))
            self._store[k][lo:hi].copy_(src, non_blocking=False)
        self._eps.append((lo, length))
        self._episode_lens = np.append(self._episode_lens, length)
        self._head = hi
        self._loaded_steps += length
        self._loaded_episodes += 1
        return True

    # ---- sampling ----
    def sample_indices(self, batch_size):
        """the reference's draws, call for call (tools/replay.py:227-229): episode ids, then window starts"""
        sequences, batch_length = self._loaded_episodes, self._length
        b_indices = np.random.randint(0, sequences, size=batch_size)
        t_indices = np.random.randint(np.zeros(batch_size) + self._min_t_sampling, self._episode_lens[b_indices] - batch_length + 1, size=batch_size)
        return b_indices, t_indices

    def gather(self, b_indices, t_indices):
        B, L = len(b_indices), self._length
        firsts = np.fromiter((self._eps[int(b)][0] for b in b_indices), dtype=np.int64, count=B)
        if self._starts_host is None or self._starts_host.numel() != B:
            self._starts_host = torch.empty(B, dtype=torch.int64).pin_memory()
            self._starts_dev = torch.empty(B, dtype=torch.int64, device=self.device)
        self._starts_host.numpy()[:] = firsts + np.asarray(t_indices, dtype=np.int64)
        self._starts_dev.copy_(self._starts_host, non_blocking=True)
        st = torch.cuda.current_stream(self.device).cuda_stream
        chunk = {}
        for k, store in self._store.items():
            dt, shape = self._meta[k]
            row = store.shape[1]
            out = torch.empty(B, L, row, device=self.device, dtype=torch.uint8)
            lib.genrl_replay_gather(st, store.data_ptr(), self._starts_dev.data_ptr(), row, B, L, out.data_ptr())
            tdt = _TORCH_DT[np.dtype(dt)]
            out = out.view(torch.bool) if tdt == torch.bool else out.view(tdt)
            chunk[k] = out.reshape((B, L) + shape)
        return chunk

    def __iter__(self):
        while True:
            yield self.gather(*self.sample_indices(self.batch_size))


def make_replay_loader(buffer, batch_size):
    """reference :310-318 wraps the buffer in a worker-less DataLoader only to get an iterator; the buffer itself is one"""
    buffer.batch_size = batch_size
    return buffer
