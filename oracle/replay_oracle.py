"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's replay sampling hot path (tools/replay.py).

  store_episode  :175-221  append an episode; while loaded_steps + length > capacity drop the oldest one
  __iter__       :223-236  b_indices = randint(0, n_episodes, B); t_indices = randint(min_t, len[b] - L + 1, B);
                           chunk[k] = stack([episode_k[b][t : t + L]])   (then torch.as_tensor(..., device))
Parity is PINNED: oracle/make_golden_replay.py imports the unmodified tools/replay.py (with a two-line stub for the absent `gym`), feeds it
and this restatement the same episodes and numpy seeds, asserts byte-identical batches and commits the reference batches' sha256 digests
(tests/golden/replay_small.json); tests/test_oracle_golden.py re-checks the digests without /root/reference.  The draws come from numpy's
global generator exactly as in the reference, so a seed fixes the batch.
"""
import numpy as np


class ReplayOracle:
    def __init__(self, length, capacity, min_t_sampling=0):
        self.length, self.capacity, self.min_t = length, capacity, min_t_sampling
        self.eps = {}                      # key -> list of arrays
        self.lens = np.array([])
        self.loaded_steps = 0

    def store_episode(self, episode):
        n = len(episode["action"])
        while self.loaded_steps + n > self.capacity:
            for k in self.eps:
                self.eps[k].pop(0)
            removed, self.lens = self.lens[0], self.lens[1:]
            self.loaded_steps -= removed
        for k, v in episode.items():
            self.eps.setdefault(k, []).append(np.asarray(v))
        self.lens = np.append(self.lens, n)
        self.loaded_steps += n

    def sample(self, batch_size):
        n_eps = len(self.lens)
        b = np.random.randint(0, n_eps, size=batch_size)
        t = np.random.randint(np.zeros(batch_size) + self.min_t, self.lens[b] - self.length + 1, size=batch_size)
        rng = np.repeat(np.expand_dims(np.arange(0, self.length), 0), batch_size, axis=0) + np.expand_dims(t, 1)
        return {k: np.stack([self.eps[k][bi][ti] for bi, ti in zip(b, rng)]) for k in self.eps}
